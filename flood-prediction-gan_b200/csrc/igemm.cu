// Implicit-GEMM convolution kernels for sm_100a: TMA-gathered operands, tcgen05.mma with TMEM accumulators.
//
//  igemm_fprop_kernel : D[pixel, k] = sum_{tap, c} A[pixel + tap, c] * B[k, (tap, c)]      (fprop, dgrad, convT)
//  igemm_wgrad_kernel : D[m, n]     = sum_{pixel}  X[pixel + xtap, m] * Y[pixel + ytap, n]  (wgrad, split over pixels)
//
// Warp roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM owner + MMA issuer (one
// elected lane), warps 2..5 = epilogue (TMEM -> registers -> global). One CTA per SM (TMEM: all 512 columns).
#include <string.h>

#include "common.cuh"
#include "host_util.h"

namespace fpg {

constexpr int kThreads = 192;
// fprop-type kernels: 8 epilogue warps, two per TMEM lane quarter (a warp may only read lanes 32 * (warp % 4) ...), each
// taking half of the tile's 16-column chunks. With 4 warps every SM scheduler held ONE epilogue warp, so the latencies of
// tcgen05.ld, the statistics shuffles and the stores were all exposed (~1.2 k clk per chunk): layers with a short K loop
// (transposed convs, PatchGAN, parity classes of the stride-2 data gradients) ran at the speed of their epilogue.
constexpr int kEpiWarps = 8;
constexpr int kFpropThreads = 64 + 32 * kEpiWarps;
constexpr int kFpropThreadsInBwd = kFpropThreads + 32;  // + the producer warp of the epilogue-operand ring
constexpr int kTmemCols = 512;

// Phase stamps of the wgrad kernels (diagnostic build: make EXTRA=-DFPG_WGRAD_TRACE): cycles from kernel entry to the
// end of setup, the accumulators complete, and the partials stored; printed for a few CTAs.
#ifdef FPG_WGRAD_TRACE
#define FPG_TRACE_DECL __shared__ long long trace_t[5];
#define FPG_TRACE_MARK(i)                                                           \
  if (threadIdx.x == 64) {                                                          \
    trace_t[i] = clock64();                                                         \
    if (i == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace_t[4]));      \
  }
#define FPG_TRACE_PRINT(name)                                                                                       \
  if (threadIdx.x == 64) {                                                                                          \
    long long t_end;                                                                                                \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));                                                       \
    printf(name " cta %d: setup %lld  mainloop %lld  epilogue %lld clk  start %lld end %lld ns\n",                  \
           static_cast<int>(blockIdx.x), trace_t[1] - trace_t[0], trace_t[2] - trace_t[1], trace_t[3] - trace_t[2], \
           trace_t[4] % 100000000ll, t_end % 100000000ll);                                                          \
  }
#else
#define FPG_TRACE_DECL
#define FPG_TRACE_MARK(i)
#define FPG_TRACE_PRINT(name)
#endif

// InstanceNorm / BatchNorm statistics from the epilogue: per (tile row-quarter, column) partial {sum, sum of squares} of
// the bf16-rounded values that are stored (a later tiny kernel reduces the rows of an image in a fixed order). Saves
// the separate pass that re-reads the whole activation.
struct StatOut {
  float* partial;        // [n_img][rows_per_img][c_total][2]; nullptr = off
  int32_t rows_per_img;  // partial rows of one image over all launches that contribute to it
  int32_t row0;          // first row of this launch inside an image
  int32_t c_total;
};

// InstanceNorm-backward reductions from a dgrad epilogue (fpg_igemm_fprop_desc.inbwd_*). The launch produces dz, the
// gradient w.r.t. the (reflect-haloed) input z of the forward convolution; z itself -- saved for the weight gradient,
// same geometry as dz -- supplies everything the reductions need, without the pre-norm tensor or its statistics:
//   mode 1, z = relu(zhat):            {sum f [z > 0], sum f z}        = {sum g', sum g' zhat}
//   mode 2, z = zprev + zhat (block):  {sum f, sum f (z - zprev)}      = {sum g,  sum g zhat}
// (f = the stored value; both sums are linear in dz and z's halo mirrors its interior, so the halo needs no fold
// first). The operand tiles are staged by a dedicated producer warp (TMA, 32-channel slices) in a shared-memory ring.
struct InBwdStat {
  int32_t mode;       // 0 = off
  int32_t has_add;    // add the skip-connection gradient to the INTERIOR outputs before they are stored / reduced
  int32_t halo;       // halo of the output tensor (tile coordinates are padded coordinates)
  int32_t h, w;       // interior extent
  int32_t add_shift;  // (halo of the skip-gradient buffer) - halo: coordinate shift of its view
  int32_t debug;      // timing experiments (FPG_INBWD_DEBUG): bit 0 = no reductions, bit 1 = no operand loads
  int32_t n_tensors;  // staged tensors per slice: z [, zprev] [, add] in this order
};
constexpr int kEStages = 3;                 // ring stages; a stage holds one 32-channel slice of each staged tensor
constexpr int kESliceBytes = 128 * 32 * 2;  // 128 pixel rows x 32 channels bf16, 64-byte swizzle

struct FpropArgs {
  int32_t chunks_per_tap;
  int32_t num_kstages;
  int32_t block_n;
  int32_t n_blocks;
  int32_t n_img, tiles_y, tiles_x;
  int32_t tile_w_log2;
  int32_t tile_h, tile_w;
  // optional second tile region (edge columns x >= x_org1 with their own tile shape and tensor map): tiles
  // [r0_tiles, r0_tiles + r1_tiles) of the persistent loop
  int32_t r0_tiles, tiles_y1, tiles_x1, tile_h1, tile_w1, tile_w1_log2, x_org1;
  int32_t m_sub;       // 128-pixel sub-tiles per CTA tile (1 or 2): each K step issues m_sub MMAs sharing one B tile
  int32_t acc_stages;  // accumulator stages in TMEM: 2 if 2 * m_sub * block_n <= 512 else 1
  int32_t act;
  int32_t stages;
  const float* bias;
  fpg_out_view out;
  StatOut stat;
  InBwdStat inbwd;
  fpg_tap taps[FPG_MAX_TAPS];
};

// Column sums over the 32 rows of a warp for 16 consecutive columns held per lane: a butterfly that halves the
// columns kept at every step (16 -> 8 -> 4 -> 2 -> 1 per lane, 15 + 1 shuffles instead of 80). Afterwards lane l holds
// the 32-row sum of column 8*b4 + 4*b3 + 2*b2 + b1 (b_k = bit k of l); lanes l and l^1 hold the same column.
__device__ __forceinline__ float warp_colsum16(float (&a)[16], int lane) {
#pragma unroll
  for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
    const bool hi = (lane & bit) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float keep = hi ? a[j + half] : a[j];
      const float send = hi ? a[j] : a[j + half];
      a[j] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  return a[0] + __shfl_xor_sync(0xffffffffu, a[0], 1);
}

__device__ __forceinline__ void stat_accumulate(const StatOut& so, const float (&f)[16], bool valid, int lane,
                                                int64_t prow, int col0, int dt) {
  float a[16], b[16];
  // statistics of the values as stored (bf16 or fp16); the element type is kernel-uniform: one branch, not 16 selects
  if (dt == 2) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = valid ? round_f16(f[i]) : 0.f;
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = valid ? __bfloat162float(__float2bfloat16(f[i])) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) b[i] = a[i] * a[i];
  const float sa = warp_colsum16(a, lane);
  const float sb = warp_colsum16(b, lane);
  if ((lane & 1) == 0) {
    const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    reinterpret_cast<float2*>(so.partial)[prow * so.c_total + col0 + col] = make_float2(sa, sb);
  }
}

__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void unpack_16x8(const uint4& u, float* f, int dt) {
  if (dt != 2) return unpack_bf16x8(u, f);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = unpack_f16x2(w[i]);
    f[2 * i] = v.x;
    f[2 * i + 1] = v.y;
  }
}
// InstanceNorm-backward variant of stat_accumulate (see InBwdStat): fr = the 16 values as stored (bf16-rounded, 0 for
// masked rows), z / zp = the forward input and, mode 2, the previous block's input at the same pixel and channels.
__device__ __forceinline__ void inbwd_accumulate(const StatOut& so, const float (&fr)[16], const float (&z)[16],
                                                 const float (&zp)[16], int mode, int lane, int64_t prow, int col0) {
  float a[16], b[16];
  if (mode == 1) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      a[k] = z[k] > 0.f ? fr[k] : 0.f;
      b[k] = fr[k] * z[k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      a[k] = fr[k];
      b[k] = fr[k] * (z[k] - zp[k]);
    }
  }
  const float sa = warp_colsum16(a, lane);
  const float sb = warp_colsum16(b, lane);
  if ((lane & 1) == 0) {
    const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    reinterpret_cast<float2*>(so.partial)[prow * so.c_total + col0 + col] = make_float2(sa, sb);
  }
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case FPG_ACT_RELU: return fmaxf(v, 0.f);
    case FPG_ACT_LEAKY: return v > 0.f ? v : 0.2f * v;
    case FPG_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// INBWD: the InstanceNorm-backward reductions of InBwdStat run in the epilogue; zmap / pmap / gmap (+ the region-1
// variants) are the views of z, zprev and the skip gradient with 32-channel boxes of the tile shapes.
template <int CBLK, bool INBWD>
__global__ void __launch_bounds__(INBWD ? kFpropThreadsInBwd : kFpropThreads, 1)
igemm_fprop_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                   const __grid_constant__ CUtensorMap amap1, const __grid_constant__ CUtensorMap zmap,
                   const __grid_constant__ CUtensorMap zmap1, const __grid_constant__ CUtensorMap pmap,
                   const __grid_constant__ CUtensorMap pmap1, const __grid_constant__ CUtensorMap gmap,
                   const __grid_constant__ CUtensorMap gmap1, const __grid_constant__ FpropArgs args) {
  constexpr int SUB = 64 / CBLK;                   // TMA sub-loads per 64-wide K stage
  constexpr uint32_t LAYOUT = swizzle_layout_type(CBLK * 2);
  constexpr uint32_t SBO = 8u * CBLK * 2u;  // 8 rows of one swizzle atom

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int BN = args.block_n;
  constexpr int MS = 1;  // 256-pixel single-CTA tiles (two MMAs per B tile) measured slower: epilogue exposed
  constexpr uint32_t A_SUB_BYTES = static_cast<uint32_t>(MS) * 128u * CBLK * 2u;  // one TMA box: MS*128 pixel rows
  constexpr uint32_t A_HALF_BYTES = 128u * CBLK * 2u;                             // rows of one 128-pixel sub-tile
  constexpr uint32_t A_STAGE_BYTES = static_cast<uint32_t>(MS) * 128u * 64u * 2u;
  const uint32_t B_SUB_BYTES = static_cast<uint32_t>(BN) * CBLK * 2u;
  const uint32_t B_STAGE_BYTES = static_cast<uint32_t>(BN) * 128u;
  const int STAGES = args.stages;
  const uint32_t ACC_COLS = static_cast<uint32_t>(MS * BN);  // TMEM columns of one accumulator stage
  const uint32_t ACC_STAGES = static_cast<uint32_t>(args.acc_stages);

  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint8_t* ering = smem_b + STAGES * B_STAGE_BYTES;  // INBWD: kEStages x estage_bytes (1 KB aligned: stage sizes are)
  const int estage_bytes = INBWD ? args.inbwd.n_tensors * kESliceBytes : 0;
  const int add_slot = (args.inbwd.mode == 2 ? 2 : 1) * kESliceBytes;  // byte offset of the skip gradient in a stage
  uint64_t* full = reinterpret_cast<uint64_t*>(ering + kEStages * estage_bytes);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* efull = tempty + 2;
  uint64_t* eempty = efull + kEStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(eempty + kEStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 32 * kEpiWarps);
    }
    for (int i = 0; i < kEStages; ++i) {
      mbar_init(&efull[i], 1);
      mbar_init(&eempty[i], kEpiWarps / 2);  // a slice is consumed by the four warps of one column half
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&amap);
    tma_prefetch_desc(&bmap);
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int r1_tiles = args.n_img * args.tiles_y1 * args.tiles_x1 * args.n_blocks;
  const int total_tiles = args.r0_tiles + r1_tiles;
  const int num_kstages = args.num_kstages;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      const uint32_t stage_bytes = A_STAGE_BYTES + B_STAGE_BYTES;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const bool r1 = tile >= args.r0_tiles;
        const int t = r1 ? tile - args.r0_tiles : tile;
        const int ntx = r1 ? args.tiles_x1 : args.tiles_x, nty = r1 ? args.tiles_y1 : args.tiles_y;
        const CUtensorMap* am = r1 ? &amap1 : &amap;
        int nb = t % args.n_blocks;
        int r = t / args.n_blocks;
        int tx = r % ntx;
        r /= ntx;
        int ty = r % nty;
        int n = r / nty;
        const int x0 = r1 ? args.x_org1 + tx * args.tile_w1 : tx * args.tile_w;
        const int y0 = ty * (r1 ? args.tile_h1 : args.tile_h);
        // the issue time of this thread is on the critical path of the pipeline: no divisions, the tap entry is
        // re-read only when the tap changes
        int tap = 0, ccol = 0, kcol = 0;
        fpg_tap tp = args.taps[0];
        const int c_per_tap = args.chunks_per_tap * CBLK;
        for (int ks = 0; ks < num_kstages; ++ks) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], stage_bytes);
          uint8_t* a_dst = smem_a + stage * A_STAGE_BYTES;
          uint8_t* b_dst = smem_b + stage * B_STAGE_BYTES;
#pragma unroll
          for (int j = 0; j < SUB; ++j) {
            tma_load_5d(am, &full[stage], a_dst + j * A_SUB_BYTES, tp.c0 + ccol, x0 + tp.dx, tp.plane, y0 + tp.dy, n);
            tma_load_2d(&bmap, &full[stage], b_dst + j * B_SUB_BYTES, kcol, nb * BN);
            kcol += CBLK;
            ccol += CBLK;
            if (ccol == c_per_tap) {
              ccol = 0;
              ++tap;
              tp = args.taps[tap & (FPG_MAX_TAPS - 1)];
            }
          }
          if (++stage == static_cast<uint32_t>(STAGES)) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t as = it % ACC_STAGES, aph = (it / ACC_STAGES) & 1;
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_COLS;
        for (int ks = 0; ks < num_kstages; ++ks) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < SUB; ++j) {
#pragma unroll
            for (int k = 0; k < CBLK / 16; ++k) {
              const uint64_t bd = make_smem_desc(b_addr + j * B_SUB_BYTES + k * 32, 0, SBO, LAYOUT);
              for (int ms = 0; ms < MS; ++ms) {
                const uint64_t ad =
                    make_smem_desc(a_addr + j * A_SUB_BYTES + ms * A_HALF_BYTES + k * 32, 0, SBO, LAYOUT);
                umma_bf16(d_tmem + ms * BN, ad, bd, idesc, (ks | j | k) != 0 ? 1u : 0u);
              }
            }
          }
          umma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == static_cast<uint32_t>(STAGES)) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[as]);  // accumulator complete
      }
    }
  } else if (INBWD && warp == 2 + kEpiWarps) {
    // ------------------------------------------------------------ producer of the epilogue-operand ring (INBWD)
    if (elect_one()) {
      if (args.inbwd.debug & 2) return;
      const uint32_t bytes = static_cast<uint32_t>(estage_bytes);
      const int slices = BN >> 6;  // 32-channel slices per column half
      const int shift = args.inbwd.add_shift;
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const bool r1 = tile >= args.r0_tiles;
        const int t = r1 ? tile - args.r0_tiles : tile;
        const int ntx = r1 ? args.tiles_x1 : args.tiles_x, nty = r1 ? args.tiles_y1 : args.tiles_y;
        int nb = t % args.n_blocks;
        int r = t / args.n_blocks;
        int tx = r % ntx;
        r /= ntx;
        int ty = r % nty;
        int n = r / nty;
        const int x0 = r1 ? args.x_org1 + tx * args.tile_w1 : tx * args.tile_w;
        const int y0 = ty * (r1 ? args.tile_h1 : args.tile_h);
        const CUtensorMap* zm = r1 ? &zmap1 : &zmap;
        const CUtensorMap* pm = r1 ? &pmap1 : &pmap;
        const CUtensorMap* gm = r1 ? &gmap1 : &gmap;
        for (int sl = 0; sl < slices; ++sl) {
          for (int half = 0; half < 2; ++half, ++g) {
            const uint32_t est = g % kEStages, eph = (g / kEStages) & 1;
            mbar_wait(&eempty[est], eph ^ 1);
            mbar_arrive_expect_tx(&efull[est], bytes);
            uint8_t* dst = ering + est * estage_bytes;
            const int c0 = nb * BN + half * (BN >> 1) + sl * 32;
            tma_load_5d(zm, &efull[est], dst, c0, x0, 0, y0, n);
            if (args.inbwd.mode == 2) tma_load_5d(pm, &efull[est], dst + kESliceBytes, c0, x0, 0, y0, n);
            if (args.inbwd.has_add)
              tma_load_5d(gm, &efull[est], dst + add_slot, c0, x0 + shift, 0, y0 + shift, n);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    // column chunks [c_begin, c_end) of this warp: first or second half of the tile's 16-column chunks
    const int n_chunks16 = BN / 16, chunk_split = (n_chunks16 + 1) / 2;
    const int c_begin = (warp - 2) < 4 ? 0 : chunk_split * 16;
    const int c_end = (warp - 2) < 4 ? chunk_split * 16 : BN;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t as = it % ACC_STAGES, aph = (it / ACC_STAGES) & 1;
      const bool r1 = tile >= args.r0_tiles;
      const int t = r1 ? tile - args.r0_tiles : tile;
      const int ntx = r1 ? args.tiles_x1 : args.tiles_x, nty = r1 ? args.tiles_y1 : args.tiles_y;
      const int tw = r1 ? args.tile_w1 : args.tile_w, th = r1 ? args.tile_h1 : args.tile_h;
      const int twl = r1 ? args.tile_w1_log2 : args.tile_w_log2;
      int nb = t % args.n_blocks;
      int r = t / args.n_blocks;
      int tx = r % ntx;
      r /= ntx;
      int ty = r % nty;
      int n = r / nty;
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      for (int ms = 0; ms < MS; ++ms) {
        const int row = ms * 128 + q * 32 + lane;
        const int ry = row >> twl;
        const int rx = row & (tw - 1);
        const int py = ty * th + ry, px = (r1 ? args.x_org1 : 0) + tx * tw + rx;
        const bool valid = (py < args.out.valid_h) && (px < args.out.valid_w);
        const int64_t off = static_cast<int64_t>(n) * args.out.stride_n +
                            static_cast<int64_t>(py * args.out.mul_y + args.out.off_y) * args.out.stride_y +
                            static_cast<int64_t>(px * args.out.mul_x + args.out.off_x) * args.out.stride_x +
                            static_cast<int64_t>(nb) * BN;
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * ACC_COLS + ms * BN;
        // statistics row of this warp: image-major, then (region, tile row, tile column, warp quarter)
        const int64_t prow = static_cast<int64_t>(n) * args.stat.rows_per_img + args.stat.row0 +
                             ((r1 ? args.tiles_y * args.tiles_x : 0) + ty * ntx + tx) * 4 + q;
        // skip-connection gradient: added to interior outputs only (INBWD)
        bool interior = false;
        if constexpr (INBWD) {
          const int iy = py - args.inbwd.halo, ix = px - args.inbwd.halo;
          interior = valid && iy >= 0 && iy < args.inbwd.h && ix >= 0 && ix < args.inbwd.w;
        }
        // one 16-column chunk: accumulator -> (+ skip gradient) -> bias / activation -> statistics -> store.
        // erow: this thread's row of the staged operand slice (INBWD), sub = which half of its 32 channels
        auto chunk = [&](int c, const uint8_t* erow, int sub, int swz) {
          uint32_t v[16];
          tmem_ld16(t_addr + c, v);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          float zv[16], zp[16];
          if constexpr (INBWD) {
            // 64-byte swizzle: the 16-byte chunk index is XORed with bits [1, 3) of the row
            const int k0 = ((2 * sub) ^ swz) << 4, k1 = ((2 * sub + 1) ^ swz) << 4;
            unpack_bf16x8(*reinterpret_cast<const uint4*>(erow + k0), zv);
            unpack_bf16x8(*reinterpret_cast<const uint4*>(erow + k1), zv + 8);
            if (args.inbwd.mode == 2) {
              unpack_bf16x8(*reinterpret_cast<const uint4*>(erow + kESliceBytes + k0), zp);
              unpack_bf16x8(*reinterpret_cast<const uint4*>(erow + kESliceBytes + k1), zp + 8);
            }
            if (args.inbwd.has_add && interior) {
              float e[16];
              unpack_bf16x8(*reinterpret_cast<const uint4*>(erow + add_slot + k0), e);
              unpack_bf16x8(*reinterpret_cast<const uint4*>(erow + add_slot + k1), e + 8);
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] += e[i];
            }
          }
          if (args.bias != nullptr) {
            const float4* bp = reinterpret_cast<const float4*>(args.bias + nb * BN + c);
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
            for (int i = 0; i < 16; ++i) f[i] = apply_act(f[i], args.act);
          }
          if (args.stat.partial != nullptr) {
            if constexpr (INBWD) {
              float fr[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) fr[i] = valid ? __bfloat162float(__float2bfloat16(f[i])) : 0.f;
              if (!(args.inbwd.debug & 1))
                inbwd_accumulate(args.stat, fr, zv, zp, args.inbwd.mode, lane, prow, nb * BN + c);
            } else {
              stat_accumulate(args.stat, f, valid, lane, prow, nb * BN + c, args.out.fp32);
            }
          }
          if (valid) {
            if (args.out.fp32 == FPG_DT_FP32) {
              store_16x32(static_cast<float*>(args.out.base) + off + c, f);
            } else {
              store_16x16(static_cast<__nv_bfloat16*>(args.out.base) + off + c, f, args.out.fp32);
            }
          }
        };
        if constexpr (!INBWD) {
          for (int c = c_begin; c < c_end; c += 16) chunk(c, nullptr, 0, 0);
        } else {
          // operand slices arrive in the order (slice 0, half 0), (slice 0, half 1), (slice 1, half 0), ...
          const int half = (warp - 2) >> 2;
          const int slices = (c_end - c_begin) >> 5;  // 32-channel slices of this half (launcher: block_n % 64 == 0)
          for (int sl = 0; sl < slices; ++sl) {
            const uint32_t g = (it * static_cast<uint32_t>(slices) + sl) * 2 + half;
            const uint32_t est = g % kEStages, eph = (g / kEStages) & 1;
            if (!(args.inbwd.debug & 2)) mbar_wait(&efull[est], eph);
            const uint8_t* erow = ering + est * estage_bytes + row * 64;
            const int swz = (row >> 1) & 3;
            chunk(c_begin + sl * 32, erow, 0, swz);
            chunk(c_begin + sl * 32 + 16, erow, 1, swz);
            __syncwarp();
            if (lane == 0 && !(args.inbwd.debug & 2)) mbar_arrive(&eempty[est]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ 2-CTA fprop
// Same implicit GEMM on a CTA PAIR (cluster of 2 = one TPC): tcgen05.mma.cta_group::2 with M = 256 pixels (128 per
// CTA) and N = block_n. Each CTA TMA-loads its own 128 pixel rows of A and only HALF of the weight tile (block_n / 2
// rows); the tensor cores read the other half from the peer's shared memory. Per CTA and K stage that is 16 KB + 16 KB
// instead of 16 KB + 32 KB: the 1-CTA kernel is bound by L2->SM operand delivery (~45 B/clk/SM), not by the MMA.
// Protocol (leader = cluster rank 0): both producers credit the LEADER's full barrier; the leader's elected thread
// issues the MMAs and multicasts the commit to both CTAs' empty / tmem-full barriers; both epilogues drain their own
// TMEM lanes and arrive on the leader's tmem-empty barrier.
template <int CBLK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFpropThreads, 1)
igemm_fprop2_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                    const __grid_constant__ FpropArgs args) {
  constexpr int SUB = 64 / CBLK;
  constexpr uint32_t A_SUB_BYTES = 128u * CBLK * 2u;
  constexpr uint32_t A_STAGE_BYTES = 128u * 64u * 2u;
  constexpr uint32_t LAYOUT = swizzle_layout_type(CBLK * 2);
  constexpr uint32_t SBO = 8u * CBLK * 2u;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int BN = args.block_n;
  const int BH = BN / 2;  // weight rows held by each CTA
  const uint32_t B_SUB_BYTES = static_cast<uint32_t>(BH) * CBLK * 2u;
  const uint32_t B_STAGE_BYTES = static_cast<uint32_t>(BH) * 128u;
  const int STAGES = args.stages;

  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 2);   // leader's copy is used: its own arrive.expect_tx + the peer's remote arrive
      mbar_init(&empty[i], 1);  // one multicast commit per use
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 2 * 32 * kEpiWarps);  // leader's copy is used: the epilogue threads of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&amap);
    tma_prefetch_desc(&bmap);
  }
  cluster_sync_all();
  if (warp == 1) tmem_alloc_2cta(tmem_slot, kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = args.n_img * args.tiles_y * args.tiles_x * args.n_blocks;  // pair tiles
  const int num_kstages = args.num_kstages;
  const int first_tile = blockIdx.x >> 1, tile_step = gridDim.x >> 1;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      const uint32_t stage_bytes = 2u * (A_STAGE_BYTES + B_STAGE_BYTES);
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        int nb = tile % args.n_blocks;
        int r = tile / args.n_blocks;
        int tx = r % args.tiles_x;
        r /= args.tiles_x;
        int ty = r % args.tiles_y;
        int n = r / args.tiles_y;
        const int x0 = tx * args.tile_w, y0 = (ty * 2 + static_cast<int>(rank)) * args.tile_h;
        int tap = 0, ccol = 0, kcol = 0;
        fpg_tap t = args.taps[0];
        const int c_per_tap = args.chunks_per_tap * CBLK;
        const int b_row = nb * BN + static_cast<int>(rank) * BH;
        for (int ks = 0; ks < num_kstages; ++ks) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(&full[stage], stage_bytes);
          } else {
            mbar_arrive_cluster(&full[stage], 0);
          }
          uint8_t* a_dst = smem_a + stage * A_STAGE_BYTES;
          uint8_t* b_dst = smem_b + stage * B_STAGE_BYTES;
#pragma unroll
          for (int j = 0; j < SUB; ++j) {
            tma_load_5d_2sm(&amap, &full[stage], a_dst + j * A_SUB_BYTES, t.c0 + ccol, x0 + t.dx, t.plane, y0 + t.dy,
                            n);
            tma_load_2d_2sm(&bmap, &full[stage], b_dst + j * B_SUB_BYTES, kcol, b_row);
            kcol += CBLK;
            ccol += CBLK;
            if (ccol == c_per_tap) {
              ccol = 0;
              ++tap;
              t = args.taps[tap & (FPG_MAX_TAPS - 1)];
            }
          }
          if (++stage == static_cast<uint32_t>(STAGES)) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      const uint32_t idesc = make_idesc_bf16(256, BN, 0, 0);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++it) {
        const uint32_t as = it & 1, aph = (it >> 1) & 1;
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int ks = 0; ks < num_kstages; ++ks) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < SUB; ++j) {
#pragma unroll
            for (int k = 0; k < CBLK / 16; ++k) {
              const uint64_t ad = make_smem_desc(a_addr + j * A_SUB_BYTES + k * 32, 0, SBO, LAYOUT);
              const uint64_t bd = make_smem_desc(b_addr + j * B_SUB_BYTES + k * 32, 0, SBO, LAYOUT);
              umma_bf16_2cta(d_tmem, ad, bd, idesc, (ks | j | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit_2cta(&empty[stage], 3);  // both CTAs' slots are free once these MMAs have read them
          if (++stage == static_cast<uint32_t>(STAGES)) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2cta(&tfull[as], 3);  // accumulators complete in both CTAs
      }
    }
  } else {
    const int q = warp & 3;
    const int n_chunks16 = BN / 16, chunk_split = (n_chunks16 + 1) / 2;  // two warps per lane quarter, see kEpiWarps
    const int c_begin = (warp - 2) < 4 ? 0 : chunk_split * 16;
    const int c_end = (warp - 2) < 4 ? chunk_split * 16 : BN;
    const int row = q * 32 + lane;
    const int ry = row >> args.tile_w_log2;
    const int rx = row & (args.tile_w - 1);
    uint32_t it = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++it) {
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      int nb = tile % args.n_blocks;
      int r = tile / args.n_blocks;
      int tx = r % args.tiles_x;
      r /= args.tiles_x;
      int ty = r % args.tiles_y;
      int n = r / args.tiles_y;
      const int py = (ty * 2 + static_cast<int>(rank)) * args.tile_h + ry, px = tx * args.tile_w + rx;
      const bool valid = (py < args.out.valid_h) && (px < args.out.valid_w);
      const int64_t off = static_cast<int64_t>(n) * args.out.stride_n +
                          static_cast<int64_t>(py * args.out.mul_y + args.out.off_y) * args.out.stride_y +
                          static_cast<int64_t>(px * args.out.mul_x + args.out.off_x) * args.out.stride_x +
                          static_cast<int64_t>(nb) * BN;
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 256;
      const int64_t prow = static_cast<int64_t>(n) * args.stat.rows_per_img + args.stat.row0 +
                           ((ty * 2 + static_cast<int>(rank)) * args.tiles_x + tx) * 4 + q;
      for (int c = c_begin; c < c_end; c += 16) {
        uint32_t v[16];
        tmem_ld16(t_addr + c, v);
        tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
        if (args.bias != nullptr) {
          const float4* bp = reinterpret_cast<const float4*>(args.bias + nb * BN + c);
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
          for (int i = 0; i < 16; ++i) f[i] = apply_act(f[i], args.act);
        }
        if (args.stat.partial != nullptr)
          stat_accumulate(args.stat, f, valid, lane, prow, nb * BN + c, args.out.fp32);
        if (valid) {
          if (args.out.fp32 == FPG_DT_FP32) {
            store_16x32(static_cast<float*>(args.out.base) + off + c, f);
          } else {
            store_16x16(static_cast<__nv_bfloat16*>(args.out.base) + off + c, f, args.out.fp32);
          }
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(&tempty[as], 0);
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ stride-2 classes
// All FOUR output-parity classes of a stride-2 data gradient / transposed-convolution forward in ONE launch.
//   out[n, 2i + pi, 2j + pj, c] = sum over the taps (r, s) of class (pi, pj), over k:  dy[n, i + ty, j + tx, k] * W[k, c, r, s]
// Launched class by class (igemm_fprop_kernel, one launch per class) these layers run at the speed of the L2 -> SM
// path, not of the tensor cores: K = taps * c_out is short (1-4 taps), so per 128-pixel tile a class moves 24 KB of
// operands per 64-wide K step for 128 x N x 64 MACs, the activation tile is fetched once per tap (9 times over the
// four classes of a 3x3 filter) and the weight tile once per tile (ncu: 8 TB/s of L2 -> SM traffic, tensor pipe 19-28 %
// active, profiles/r02). Here a CTA keeps EVERY tap's weights resident in shared memory for its lifetime and walks the
// distinct input shifts (ty, tx) -- four for a 3x3 filter -- loading each shifted activation tile once and issuing
// the MMAs of every class that has a tap with that shift into that class's accumulator (4 x N <= 256 TMEM columns,
// two accumulator stages). 3.4x less operand traffic for the 3x3 layers.
constexpr int kS2MaxShifts = 9;
constexpr int kS2MaxOps = 4 * 4 * 8;  // (taps over the four classes) x K chunks: 16 taps x 8 chunks of 64 channels
struct S2Args {
  int32_t n_img, tiles_y, tiles_x, tile_w_log2, tile_h, tile_w;  // tiling of the class grid (= the dy pixel grid)
  int32_t block_n;   // N = c_in <= 64
  int32_t kchunks;   // c_out / 64
  int32_t stages;    // activation ring
  int32_t n_shifts;
  int32_t line_store;       // BN == 64: lanes of a group of four write one pixel's 128-byte line together
  int32_t cls_taps[4];      // taps of each class
  int32_t cls_slot0[4];     // first resident weight slot of each class (slot = one tap x one 64-wide K chunk)
  struct Shift {
    int16_t ty, tx;
    int16_t n_users, pad;
    struct { int16_t cls, slot; } u[4];  // slot: resident slot of (class, tap) at K chunk 0
  } shift[kS2MaxShifts];
  int32_t off_y[4], off_x[4];  // output parity of each class
  int32_t stat_row0[4];
  fpg_out_view out;            // mul 2, offsets per class above
  StatOut stat;
};

__global__ void __launch_bounds__(kFpropThreads, 1)
igemm_s2cls_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap0,
                   const __grid_constant__ CUtensorMap bmap1, const __grid_constant__ CUtensorMap bmap2,
                   const __grid_constant__ CUtensorMap bmap3, const __grid_constant__ S2Args args) {
  constexpr uint32_t LAYOUT = swizzle_layout_type(128);
  constexpr uint32_t SBO = 8u * 128u;
  constexpr uint32_t A_STAGE_BYTES = 128u * 64u * 2u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int BN = args.block_n;
  const uint32_t SLOT_BYTES = static_cast<uint32_t>(BN) * 128u;
  const int STAGES = args.stages;
  int n_slots = 0;
  for (int q = 0; q < 4; ++q) n_slots += args.cls_taps[q] * args.kchunks;
  uint8_t* smem_b = smem;                                   // resident weights
  uint8_t* smem_a = smem + ((n_slots * SLOT_BYTES + 1023u) & ~1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_a + STAGES * A_STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* bfull = tempty + 2;
  uint64_t* op_bdesc = bfull + 1;  // tabulated MMA groups: kS2MaxOps weight descriptors + info words
  uint32_t* op_info = reinterpret_cast<uint32_t*>(op_bdesc + kS2MaxOps);
  uint32_t* tmem_slot = op_info + kS2MaxOps;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 32 * kEpiWarps);
    }
    mbar_init(bfull, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&amap);
    tma_prefetch_desc(&bmap0);
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total_tiles = args.n_img * args.tiles_y * args.tiles_x;
  const int KC = args.kchunks;

  if (warp == 0) {
    if (elect_one()) {
      // every tap's weights, once: slot (class q, tap t, K chunk kc) <- columns [(t * KC + kc) * 64, +64) of class q's matrix
      mbar_arrive_expect_tx(bfull, static_cast<uint32_t>(n_slots) * SLOT_BYTES);
      const CUtensorMap* bm[4] = {&bmap0, &bmap1, &bmap2, &bmap3};
      for (int q = 0; q < 4; ++q)
        for (int t = 0; t < args.cls_taps[q]; ++t)
          for (int kc = 0; kc < KC; ++kc)
            tma_load_2d(bm[q], bfull, smem_b + (args.cls_slot0[q] + t * KC + kc) * SLOT_BYTES, (t * KC + kc) * 64, 0);
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int r = tile;
        const int tx = r % args.tiles_x;
        r /= args.tiles_x;
        const int ty = r % args.tiles_y;
        const int n = r / args.tiles_y;
        const int x0 = tx * args.tile_w, y0 = ty * args.tile_h;
        for (int si = 0; si < args.n_shifts; ++si) {
          const int sx = args.shift[si].tx, sy = args.shift[si].ty;
          for (int kc = 0; kc < KC; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], A_STAGE_BYTES);
            tma_load_5d(&amap, &full[stage], smem_a + stage * A_STAGE_BYTES, kc * 64, x0 + sx, 0, y0 + sy, n);
            if (++stage == static_cast<uint32_t>(STAGES)) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      mbar_wait(bfull, 0);
      tc_fence_after();
      const uint32_t b_base = smem_u32(smem_b);
      // The sequence of MMA groups is the same for every tile: it is tabulated once in shared memory (the single
      // issuing thread is on the critical path; walking the shift / user tables in the kernel parameters per MMA cost
      // more than the MMAs). Entry: weight descriptor, accumulator column, flags {first of its class, last of its stage}.
      uint32_t n_ops = 0;
      {
        uint32_t started = 0;
        for (int si = 0; si < args.n_shifts; ++si) {
          const int nu = args.shift[si].n_users;
          for (int kc = 0; kc < KC; ++kc)
            for (int u = 0; u < nu; ++u, ++n_ops) {
              const int cls = args.shift[si].u[u].cls;
              op_bdesc[n_ops] = make_smem_desc(b_base + (args.shift[si].u[u].slot + kc) * SLOT_BYTES, 0, SBO, LAYOUT);
              op_info[n_ops] = static_cast<uint32_t>(cls * BN) | ((kc == 0 && !((started >> cls) & 1u)) ? 1u << 16 : 0u) |
                               (u == nu - 1 ? 1u << 17 : 0u);
            }
          for (int u = 0; u < nu; ++u) started |= 1u << args.shift[si].u[u].cls;
        }
      }
      const uint64_t a_hi = make_smem_desc(0, 0, SBO, LAYOUT);  // descriptor without the start address
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t as = it & 1, aph = (it >> 1) & 1;
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        bool need_stage = true;
        uint64_t ad = 0;
        for (uint32_t j = 0; j < n_ops; ++j) {
          if (need_stage) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            ad = a_hi | static_cast<uint64_t>((smem_u32(smem_a + stage * A_STAGE_BYTES) >> 4) & 0x3FFFu);
            need_stage = false;
          }
          const uint64_t bd = op_bdesc[j];
          const uint32_t info = op_info[j];
          const uint32_t d_col = d_tmem + (info & 0xFFFFu);
          umma_bf16(d_col, ad, bd, idesc, (info >> 16) & 1u ? 0u : 1u);
          umma_bf16(d_col, ad + 2, bd + 2, idesc, 1u);  // + 32 bytes = + 2 in the 16-byte address field
          umma_bf16(d_col, ad + 4, bd + 4, idesc, 1u);
          umma_bf16(d_col, ad + 6, bd + 6, idesc, 1u);
          if ((info >> 17) & 1u) {
            umma_commit(&empty[stage]);
            if (++stage == static_cast<uint32_t>(STAGES)) {
              stage = 0;
              phase ^= 1;
            }
            need_stage = true;
          }
        }
        umma_commit(&tfull[as]);
      }
    }
  } else {
    const int q = warp & 3;
    const int cls0 = (warp - 2) < 4 ? 0 : 2;  // two classes per column half of the epilogue warps
    const int row = q * 32 + lane;
    const int ry = row >> args.tile_w_log2;
    const int rx = row & (args.tile_w - 1);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      int r = tile;
      const int tx = r % args.tiles_x;
      r /= args.tiles_x;
      const int ty = r % args.tiles_y;
      const int n = r / args.tiles_y;
      const int py = ty * args.tile_h + ry, px = tx * args.tile_w + rx;
      const bool valid = (py < args.out.valid_h) && (px < args.out.valid_w);
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      for (int cls = cls0; cls < cls0 + 2; ++cls) {
        const int64_t off = static_cast<int64_t>(n) * args.out.stride_n +
                            static_cast<int64_t>(py * 2 + args.off_y[cls]) * args.out.stride_y +
                            static_cast<int64_t>(px * 2 + args.off_x[cls]) * args.out.stride_x;
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 256 + cls * BN;
        const int64_t prow = static_cast<int64_t>(n) * args.stat.rows_per_img + args.stat_row0[cls] +
                             (ty * args.tiles_x + tx) * 4 + q;
        if (BN == 64 && args.line_store && args.out.fp32 != FPG_DT_FP32) {
          // 64 produced channels = 128 bytes per pixel = four 32-byte chunks. Row-per-thread stores touch 32 different
          // lines per instruction (ncu: the epilogue warps spent a quarter of their time on store back-pressure): the
          // four chunks of four neighbouring pixels are transposed among groups of four lanes (two shuffle butterflies),
          // so that lanes 4g .. 4g+3 then write ONE pixel's 128-byte line together: 8 full lines per store instruction.
          uint32_t w[4][8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t v[16];
            tmem_ld16(t_addr + 16 * j, v);
            tmem_ld_wait();
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
            if (args.stat.partial != nullptr) stat_accumulate(args.stat, f, valid, lane, prow, 16 * j, args.out.fp32);
            if (args.out.fp32 == 2) {
#pragma unroll
              for (int i = 0; i < 8; ++i) w[j][i] = pack_f16x2(f[2 * i], f[2 * i + 1]);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) w[j][i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
            }
          }
          // 4 x 4 transpose of 32-byte blocks within lane groups of four: afterwards w[i] = chunk (lane & 3) of the
          // pixel of lane (lane & ~3) + i
#pragma unroll
          for (int bit = 1; bit <= 2; bit <<= 1) {
            const bool hi = (lane & bit) != 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j & bit) continue;  // pairs (j, j + bit)
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint32_t send = hi ? w[j][i] : w[j + bit][i];
                const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, bit);
                if (hi) {
                  w[j][i] = recv;
                } else {
                  w[j + bit][i] = recv;
                }
              }
            }
          }
          const int k = lane & 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int prow_i = q * 32 + (lane & ~3) + i;  // tile row of the pixel this store belongs to
            const int py_i = ty * args.tile_h + (prow_i >> args.tile_w_log2);
            const int px_i = tx * args.tile_w + (prow_i & (args.tile_w - 1));
            if (py_i < args.out.valid_h && px_i < args.out.valid_w) {
              const int64_t off_i = static_cast<int64_t>(n) * args.out.stride_n +
                                    static_cast<int64_t>(py_i * 2 + args.off_y[cls]) * args.out.stride_y +
                                    static_cast<int64_t>(px_i * 2 + args.off_x[cls]) * args.out.stride_x + 16 * k;
              __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(args.out.base) + off_i;
              if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
                st_global_256(dst, w[i]);
              } else {
                reinterpret_cast<uint4*>(dst)[0] = make_uint4(w[i][0], w[i][1], w[i][2], w[i][3]);
                reinterpret_cast<uint4*>(dst)[1] = make_uint4(w[i][4], w[i][5], w[i][6], w[i][7]);
              }
            }
          }
          continue;
        }
        for (int c = 0; c < BN; c += 16) {
          uint32_t v[16];
          tmem_ld16(t_addr + c, v);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          if (args.stat.partial != nullptr) stat_accumulate(args.stat, f, valid, lane, prow, c, args.out.fp32);
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
      mbar_arrive(&tempty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ wgrad
struct WgradArgs {
  int32_t x_ca, y_ca, x_atoms, y_atoms;
  int32_t x_groups, y_groups, x_taps_mode, y_taps_mode, x_ntaps, y_ntaps;
  int32_t n_img, kt_y, kt_x, tile_h, tile_w;
  int32_t splits, stages;
  int32_t x_shift_atoms, y_shift_atoms, y_shifts, y_sets;
  uint32_t x_box_bytes, y_box_bytes;    // bytes one TMA box delivers
  uint32_t x_atom_stride, y_atom_stride;  // shared-memory pitch between separately loaded boxes (1 KB multiple)
  uint32_t x_stage_bytes, y_stage_bytes;
  float* ws;
  fpg_tap x_taps[FPG_MAX_TAPS];
  fpg_tap y_taps[FPG_MAX_TAPS];
};

// Loop-invariant pieces of the MMA descriptors of one wgrad CTA. A single thread issues every MMA and the tensor pipe
// idles whenever that thread needs more cycles per MMA than the MMA takes (ncu: 41 % tensor-active with ~15 uniform
// instructions per N = 128 MMA), so the per-stage issue routine is fully unrolled per (shift groups, M halves).
struct WgradIssue {
  uint32_t tmem;
  uint64_t x_hi, y_hi;                 // descriptors without the start-address field
  uint32_t x_k16, y_k16, x_h16, y_g16;  // 16-byte units: K step (16 pixel rows), second M half, MMA group stride
  uint32_t n, idesc;
  uint64_t* full;
  uint64_t* empty;
  uint32_t x_base, y_base, x_stage, y_stage, stages;
};

template <int YS, int MSUB>
__device__ __forceinline__ void wgrad_mma_loop(const WgradIssue& is, int n_kt) {
  uint32_t stage = 0, phase = 0, acc = 0;
  for (int kt = 0; kt < n_kt; ++kt) {
    mbar_wait(&is.full[stage], phase);
    tc_fence_after();
    const uint32_t x_lo = ((is.x_base + stage * is.x_stage) >> 4) & 0x3FFFu;
    const uint32_t y_lo = ((is.y_base + stage * is.y_stage) >> 4) & 0x3FFFu;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t xd0 = is.x_hi | static_cast<uint64_t>(x_lo + k * is.x_k16);
      const uint64_t xd1 = is.x_hi | static_cast<uint64_t>(x_lo + is.x_h16 + k * is.x_k16);
#pragma unroll
      for (int g = 0; g < YS; ++g) {
        const uint64_t yd = is.y_hi | static_cast<uint64_t>(y_lo + k * is.y_k16 + g * is.y_g16);
        umma_bf16(is.tmem + (g * MSUB) * is.n, xd0, yd, is.idesc, k == 0 ? acc : 1u);
        if (MSUB == 2) umma_bf16(is.tmem + (g * MSUB + 1) * is.n, xd1, yd, is.idesc, k == 0 ? acc : 1u);
      }
    }
    acc = 1u;
    umma_commit(&is.empty[stage]);
    if (++stage == is.stages) {
      stage = 0;
      phase ^= 1;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
igemm_wgrad_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap ymap,
                   const __grid_constant__ WgradArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  __shared__ int4 boxtab[32];  // producer thread only: TMA coordinates of the X boxes [0,16) and Y boxes [16,32)
  const int M = args.x_atoms * args.x_ca;
  const int N = args.y_atoms * args.y_ca;
  const int YS = args.y_shifts * args.y_sets;  // MMA groups per stage, each with its own accumulator columns
  const uint32_t X_STAGE_BYTES = args.x_stage_bytes;
  const uint32_t Y_STAGE_BYTES = args.y_stage_bytes;
  const int STAGES = args.stages;

  uint8_t* smem_x = smem;
  uint8_t* smem_y = smem + STAGES * X_STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_y + STAGES * Y_STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  FPG_TRACE_DECL
  FPG_TRACE_MARK(0)

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&xmap);
    tma_prefetch_desc(&ymap);
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  FPG_TRACE_MARK(1)

  // work decomposition: blockIdx.x = (split, xi, yi)
  const int NX = args.x_taps_mode ? args.x_groups : args.x_groups * args.x_ntaps;
  const int NY = args.y_taps_mode ? (args.y_groups + args.y_sets - 1) / args.y_sets : args.y_groups * args.y_ntaps;
  const int items = NX * NY;
  const int item = blockIdx.x % items;
  const int split = blockIdx.x / items;
  const int xi = item / NY, yi = item % NY;
  const int total_kt = args.n_img * args.kt_y * args.kt_x;
  const int kt_begin = static_cast<int>(static_cast<int64_t>(total_kt) * split / args.splits);
  const int kt_end = static_cast<int>(static_cast<int64_t>(total_kt) * (split + 1) / args.splits);

  if (warp == 0) {
    if (elect_one()) {
      // operand tile -> (tap, channel offset) per separately loaded box
      const int xg = args.x_taps_mode ? xi : xi % args.x_groups;
      const int xt = args.x_taps_mode ? 0 : xi / args.x_groups;
      const int yg = args.y_taps_mode ? yi : yi % args.y_groups;
      const int yt = args.y_taps_mode ? 0 : yi / args.y_groups;
      const int x_boxes = args.x_shift_atoms ? 1 : args.x_atoms;
      const int y_boxes = args.y_shift_atoms ? args.y_sets : args.y_atoms;
      // per-box TMA coordinates (channel, dx, plane, dy) are loop invariant: with three 64 KB stages the producer's
      // issue time is on the critical path (slot period = issue + L2 latency + MMA), so they are tabulated once
      // (shared memory, up to 16 boxes per operand), nothing but the tile origin is computed per stage and the k-tile
      // counter is advanced without divisions
      for (int a = 0; a < x_boxes; ++a) {
        int tap = 0, coff = 0;
        if (args.x_taps_mode) {
          tap = xg * args.x_atoms + a;
          if (tap >= args.x_ntaps) tap = 0;
        } else {
          tap = xt;
          coff = (xg * args.x_atoms + a) * args.x_ca;
        }
        const fpg_tap t = args.x_taps[tap];
        boxtab[a] = make_int4(t.c0 + coff, t.dx, t.plane, t.dy);
      }
      for (int a = 0; a < y_boxes; ++a) {
        int tap = 0, coff = 0;
        if (args.y_shift_atoms) {  // box a = tap group yg * y_sets + a (its first tap; the others are pixel shifts)
          tap = (yg * args.y_sets + a) * args.y_atoms;
          if (tap >= args.y_ntaps) tap = 0;
        } else if (args.y_taps_mode) {
          tap = yg * args.y_atoms + a;
          if (tap >= args.y_ntaps) tap = 0;
        } else {
          tap = yt;
          coff = (yg * args.y_atoms + a) * args.y_ca;
        }
        const fpg_tap t = args.y_taps[tap];
        boxtab[16 + a] = make_int4(t.c0 + coff, t.dx, t.plane, t.dy);
      }
      uint32_t stage = 0, phase = 0;
      const uint32_t stage_bytes = x_boxes * args.x_box_bytes + y_boxes * args.y_box_bytes;
      int kx = kt_begin % args.kt_x;
      int ky = (kt_begin / args.kt_x) % args.kt_y;
      int n = kt_begin / (args.kt_x * args.kt_y);
      for (int kt = kt_begin; kt < kt_end; ++kt) {
        const int x0 = kx * args.tile_w, y0 = ky * args.tile_h;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], stage_bytes);
        uint8_t* x_dst = smem_x + stage * X_STAGE_BYTES;
        uint8_t* y_dst = smem_y + stage * Y_STAGE_BYTES;
        for (int a = 0; a < x_boxes; ++a) {
          const int4 b = boxtab[a];
          tma_load_5d(&xmap, &full[stage], x_dst + a * args.x_atom_stride, b.x, x0 + b.y, b.z, y0 + b.w, n);
        }
        for (int a = 0; a < y_boxes; ++a) {
          const int4 b = boxtab[16 + a];
          tma_load_5d(&ymap, &full[stage], y_dst + a * args.y_atom_stride, b.x, x0 + b.y, b.z, y0 + b.w, n);
        }
        if (++stage == static_cast<uint32_t>(STAGES)) {
          stage = 0;
          phase ^= 1;
        }
        if (++kx == args.kt_x) {
          kx = 0;
          if (++ky == args.kt_y) {
            ky = 0;
            ++n;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // both operands MN-major (pixel rows, channel-contiguous). M == 256 is issued as two M=128 MMAs that share the Y
      // tile; y_shifts > 1 issues one MMA group per pixel shift of the Y tile. Accumulator (shift g, half ms) lives
      // at TMEM columns (g * msub + ms) * N.
      const int MI = M > 128 ? 128 : M;
      const int msub = M / MI;
      const uint32_t idesc = make_idesc_bf16(MI, N, 1, 1);
      const uint32_t x_layout = swizzle_layout_type(args.x_ca * 2), y_layout = swizzle_layout_type(args.y_ca * 2);
      const uint32_t x_sbo = 8u * args.x_ca * 2u, y_sbo = 8u * args.y_ca * 2u;  // 8 pixel rows
      const uint32_t x_kstep = 16u * args.x_ca * 2u, y_kstep = 16u * args.y_ca * 2u;  // 16 pixel rows per MMA
      // atom stride seen by the tensor core: one pixel row for shift atoms, the box pitch otherwise
      const uint32_t x_lbo = args.x_shift_atoms ? args.x_ca * 2u : args.x_atom_stride;
      const uint32_t y_lbo = args.y_shift_atoms ? args.y_ca * 2u : args.y_atom_stride;
      const uint32_t x_half = args.x_shift_atoms ? 0u : 2u * args.x_atom_stride;  // second M = 128 half (4 channel atoms)
      // descriptors = constant high part | (start address >> 4): the issue loop only adds precomputed 16-byte offsets
      // (a single thread issues every MMA; ~25 ALU instructions per MMA made the N = 128 plan issue bound)
      const uint64_t x_hi = make_smem_desc(0, x_lbo, x_sbo, x_layout);
      const uint64_t y_hi = make_smem_desc(0, y_lbo, y_sbo, y_layout);
      // MMA group stride: the next box (tap sets) or one pixel row (shift groups)
      const uint32_t x_k16 = x_kstep >> 4, y_k16 = y_kstep >> 4, x_h16 = x_half >> 4,
                     y_g16 = (args.y_sets > 1 ? args.y_atom_stride : args.y_ca * 2u) >> 4;
      WgradIssue is;
      is.tmem = tmem_base;
      is.x_hi = x_hi;
      is.y_hi = y_hi;
      is.x_k16 = x_k16;
      is.y_k16 = y_k16;
      is.x_h16 = x_h16;
      is.y_g16 = y_g16;
      is.n = static_cast<uint32_t>(N);
      is.idesc = idesc;
      is.full = full;
      is.empty = empty;
      is.x_base = smem_u32(smem_x);
      is.y_base = smem_u32(smem_y);
      is.x_stage = X_STAGE_BYTES;
      is.y_stage = Y_STAGE_BYTES;
      is.stages = static_cast<uint32_t>(STAGES);
      const int n_kt = kt_end - kt_begin;
      switch ((YS - 1) * 2 + (msub - 1)) {  // (MMA groups, M halves) -> fully unrolled main loop
        case 0: wgrad_mma_loop<1, 1>(is, n_kt); break;
        case 1: wgrad_mma_loop<1, 2>(is, n_kt); break;
        case 2: wgrad_mma_loop<2, 1>(is, n_kt); break;
        case 3: wgrad_mma_loop<2, 2>(is, n_kt); break;
        case 4: wgrad_mma_loop<3, 1>(is, n_kt); break;
        case 6: wgrad_mma_loop<4, 1>(is, n_kt); break;
        default: __trap();  // never planned: the accumulators would not fit TMEM
      }
      umma_commit(tfull);
    }
  } else {
    const int q = warp & 3;
    // M >= 128: row = 32q + lane (+128 for the second sub-tile). M == 64: rows 16q..16q+15 are lanes 0..15 of quarter q
    const int msub = M > 128 ? 2 : 1;
    const bool row_valid = (M >= 128) || (lane < 16);
    const int NT = YS * N;  // workspace row length
    if (kt_end > kt_begin) {
      mbar_wait(tfull, 0);
      tc_fence_after();
    }
    FPG_TRACE_MARK(2)
    // Workspace layout (decoded by wgrad_reduce_kernel): an item is [row block of RB rows][16-column chunk][4][RB][4]
    // floats, so one store instruction of a warp writes RB x 16 contiguous bytes (row-major rows of NT floats made
    // every lane hit its own 128-byte line: 8x the L2 write requests for the same bytes).
    const int RB = (M >= 128) ? 32 : 16;
    float* item_ws = args.ws + (static_cast<int64_t>(split) * items + item) * M * NT;
    for (int ms = 0; ms < msub; ++ms) {
      const int rowblk = (M >= 128) ? ms * 4 + q : q;
      for (int g = 0; g < YS; ++g) {
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (g * msub + ms) * N;
        for (int c = 0; c < N; c += 16) {
          uint32_t v[16];
          if (kt_end > kt_begin) {
            tmem_ld16(t_addr + c, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0u;
          }
          if (row_valid) {
            float4* d4 = reinterpret_cast<float4*>(item_ws + (static_cast<int64_t>(rowblk) * (NT >> 4) + ((g * N + c) >> 4)) * (RB * 16)) + lane;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              d4[i * RB] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                       __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          }
        }
      }
    }
    FPG_TRACE_MARK(3)
    FPG_TRACE_PRINT("wgrad")
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ 2-CTA wgrad
// Weight gradient of the wide layers (c_out = 256, c_in = 256: the residual 3x3 convs) on a CTA PAIR with
// tcgen05.mma.cta_group::2. Pair tile: D[256 output channels, 2 taps x 256 input channels]. CTA r loads only ITS half of
// both operands -- dy channels [128 r, +128) (the M rows it owns) and input channels [128 r, +128) of both taps (its
// half of every N = 256 MMA) -- so a stage is 48 KB per SM for 8 MMAs of 128 cycles: 47 B/clk of L2->SM traffic
// instead of 62 (the 1-CTA M = 256 / N = 256 plan is L2-bound at 0.67 tensor-active) and 8 KB of shared-memory operand
// reads per MMA instead of 12. Each CTA's TMEM holds its 128 rows x (2 x 256) fp32 columns: all 512 columns, one use.
// Barrier protocol as in igemm_fprop2_kernel (leader = cluster rank 0 issues, commits are multicast to both CTAs).
struct Wgrad2Args {
  int32_t n_img, kt_y, kt_x, tile_h, tile_w;
  int32_t splits, last_splits, stages, ntaps, items;
  float* ws;
  fpg_tap y_taps[FPG_MAX_TAPS];
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
igemm_wgrad2_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap ymap,
                    const __grid_constant__ Wgrad2Args args) {
  constexpr uint32_t ATOM = 8192;              // 64 pixel rows x 64 channels bf16
  constexpr uint32_t X_STAGE = 2 * ATOM;       // this CTA's 128 dy channels
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // clusters [0, pairs * splits): (tap pair, split); with an odd tap count the last tap is an item of its own with
  // last_splits splits. Its stage holds one tap (32 KB instead of 48), so the same ring memory gives it 1.5x the stages:
  // a single-tap k-tile is only 4 MMAs (512 clk) and needs the deeper look-ahead to cover the L2 latency.
  const int cluster = blockIdx.x >> 1;
  const int pairs = args.ntaps >> 1;
  const bool single = cluster >= pairs * args.splits;
  const int nt = single ? 1 : 2;  // taps of this item
  const uint32_t Y_STAGE = static_cast<uint32_t>(nt) * 2u * ATOM;
  const int STAGES = single ? args.stages * 3 / 2 : args.stages;
  uint8_t* smem_x = smem;
  uint8_t* smem_y = smem + STAGES * X_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + args.stages * (X_STAGE + 4 * ATOM));
  uint64_t* empty = full + 8;
  uint64_t* tfull = empty + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  FPG_TRACE_DECL
  FPG_TRACE_MARK(0)

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 2);   // leader's copy: its own arrive.expect_tx + the peer's remote arrive
      mbar_init(&empty[i], 1);  // one multicast commit per use
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&xmap);
    tma_prefetch_desc(&ymap);
  }
  cluster_sync_all();
  if (warp == 1) tmem_alloc_2cta(tmem_slot, kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  FPG_TRACE_MARK(1)

  const int item = single ? pairs : cluster % pairs;
  const int split = single ? cluster - pairs * args.splits : cluster / pairs;
  const int my_splits = single ? args.last_splits : args.splits;
  const int total_kt = args.n_img * args.kt_y * args.kt_x;
  const int kt_begin = static_cast<int>(static_cast<int64_t>(total_kt) * split / my_splits);
  const int kt_end = static_cast<int>(static_cast<int64_t>(total_kt) * (split + 1) / my_splits);

  if (warp == 0) {
    if (elect_one()) {
      fpg_tap tp[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) tp[j] = args.y_taps[item * 2 + j < args.ntaps ? item * 2 + j : 0];
      const uint32_t stage_tx = 2u * (X_STAGE + Y_STAGE);  // both CTAs' bytes
      const int c0 = static_cast<int>(rank) * 128;
      uint32_t stage = 0, phase = 0;
      int kx = kt_begin % args.kt_x;
      int ky = (kt_begin / args.kt_x) % args.kt_y;
      int n = kt_begin / (args.kt_x * args.kt_y);
      for (int kt = kt_begin; kt < kt_end; ++kt) {
        const int x0 = kx * args.tile_w, y0 = ky * args.tile_h;
        mbar_wait(&empty[stage], phase ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&full[stage], stage_tx);
        } else {
          mbar_arrive_cluster(&full[stage], 0);
        }
        uint8_t* x_dst = smem_x + stage * X_STAGE;
        uint8_t* y_dst = smem_y + stage * Y_STAGE;
#pragma unroll
        for (int a = 0; a < 2; ++a) tma_load_5d_2sm(&xmap, &full[stage], x_dst + a * ATOM, c0 + a * 64, x0, 0, y0, n);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (j < nt) {
#pragma unroll
            for (int a = 0; a < 2; ++a)
              tma_load_5d_2sm(&ymap, &full[stage], y_dst + (j * 2 + a) * ATOM, tp[j].c0 + c0 + a * 64, x0 + tp[j].dx,
                              tp[j].plane, y0 + tp[j].dy, n);
          }
        if (++stage == static_cast<uint32_t>(STAGES)) {
          stage = 0;
          phase ^= 1;
        }
        if (++kx == args.kt_x) {
          kx = 0;
          if (++ky == args.kt_y) {
            ky = 0;
            ++n;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      const uint32_t idesc = make_idesc_bf16(256, 256, 1, 1);
      const uint64_t hi = make_smem_desc(0, ATOM, 1024, 2);  // MN-major, 128B swizzle, atom stride 8 KB, 8-row groups
      uint32_t stage = 0, phase = 0, acc = 0;
      for (int kt = kt_begin; kt < kt_end; ++kt) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t x_lo = (smem_u32(smem_x + stage * X_STAGE) >> 4) & 0x3FFFu;
        const uint32_t y_lo = (smem_u32(smem_y + stage * Y_STAGE) >> 4) & 0x3FFFu;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t xd = hi | static_cast<uint64_t>(x_lo + k * 128);  // 16 pixel rows = 2 KB = 128 x 16 B
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (j < nt) {
              const uint64_t yd = hi | static_cast<uint64_t>(y_lo + j * (2 * ATOM >> 4) + k * 128);
              umma_bf16_2cta(tmem_base + j * 256, xd, yd, idesc, k == 0 ? acc : 1u);
            }
          }
        }
        acc = 1u;
        umma_commit_2cta(&empty[stage], 3);
        if (++stage == static_cast<uint32_t>(STAGES)) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit_2cta(tfull, 3);
    }
  } else {
    // each CTA drains its own 128 rows (output channels 128 r + ...) x 512 columns = (tap j, input channel)
    const int q = warp & 3;
    if (kt_end > kt_begin) {
      mbar_wait(tfull, 0);
      tc_fence_after();
    }
    FPG_TRACE_MARK(2)
    // workspace layout as in igemm_wgrad_kernel: [row block of 32][16-column chunk][4][32 lanes][4] floats
    const int rowblk = static_cast<int>(rank) * 4 + q;
    float* item_ws = args.ws + (static_cast<int64_t>(split) * args.items + item) * 256 * 512;
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int c = 0; c < nt * 256; c += 16) {
      uint32_t v[16];
      if (kt_end > kt_begin) {
        tmem_ld16(t_addr + c, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0u;
      }
      float4* d4 = reinterpret_cast<float4*>(item_ws + (static_cast<int64_t>(rowblk) * 32 + (c >> 4)) * 512) + lane;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        d4[i * 32] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                 __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
    }
    FPG_TRACE_MARK(3)
    FPG_TRACE_PRINT("wgrad2")
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem_base, kTmemCols);
}

static int log2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}

}  // namespace fpg

using namespace fpg;

extern "C" int fpg_igemm_fprop_launch(const fpg_igemm_fprop_desc* d, void* stream) {
  FPG_REQUIRE(d != nullptr, "null descriptor");
  FPG_REQUIRE(d->cblk == 16 || d->cblk == 32 || d->cblk == 64, "cblk %d", d->cblk);
  FPG_REQUIRE(d->block_n >= 16 && d->block_n <= 256 && d->block_n % 16 == 0, "block_n %d", d->block_n);
  const int m_sub = d->tile_h * d->tile_w / 128;
  FPG_REQUIRE(m_sub == 1 && d->tile_h * d->tile_w == 128 * m_sub && log2_exact(d->tile_w) >= 0,
              "tile %dx%d", d->tile_h, d->tile_w);
  FPG_REQUIRE(m_sub * d->block_n <= 512, "accumulator does not fit TMEM");
  const int sub_per_stage = 64 / d->cblk;
  FPG_REQUIRE(d->num_sub > 0 && d->num_sub % sub_per_stage == 0, "num_sub %d", d->num_sub);
  FPG_REQUIRE(d->c_per_tap % d->cblk == 0, "c_per_tap %d", d->c_per_tap);
  FPG_REQUIRE(d->num_taps <= FPG_MAX_TAPS && d->num_taps * (d->c_per_tap / d->cblk) == d->num_sub, "taps %d",
              d->num_taps);
  FPG_REQUIRE(d->stages >= 2 && d->stages <= 8, "stages %d", d->stages);
  CUtensorMap amap, bmap;
  int rc = encode_tmap(&d->a, &amap);
  if (rc) return rc;
  rc = encode_tmap(&d->b, &bmap);
  if (rc) return rc;

  FpropArgs args;
  args.chunks_per_tap = d->c_per_tap / d->cblk;
  args.num_kstages = d->num_sub / sub_per_stage;
  args.block_n = d->block_n;
  args.n_blocks = d->n_blocks;
  args.n_img = d->n_img;
  args.tiles_y = d->tiles_y;
  args.tiles_x = d->tiles_x;
  args.tile_w_log2 = log2_exact(d->tile_w);
  args.tile_h = d->tile_h;
  args.tile_w = d->tile_w;
  args.r0_tiles = d->n_img * d->tiles_y * d->tiles_x * d->n_blocks;
  const bool two_regions = d->tiles_x1 > 0 && d->tiles_y1 > 0;
  args.tiles_y1 = two_regions ? d->tiles_y1 : 0;
  args.tiles_x1 = two_regions ? d->tiles_x1 : 0;
  args.tile_h1 = two_regions ? d->tile_h1 : 1;
  args.tile_w1 = two_regions ? d->tile_w1 : 1;
  args.tile_w1_log2 = two_regions ? log2_exact(d->tile_w1) : 0;
  args.x_org1 = two_regions ? d->x_org1 : 0;
  CUtensorMap amap1 = amap;
  if (two_regions) {
    FPG_REQUIRE(!d->cta_pair && d->tile_h1 * d->tile_w1 == 128 && log2_exact(d->tile_w1) >= 0, "second tile region");
    rc = encode_tmap(&d->a1, &amap1);
    if (rc) return rc;
  }
  args.m_sub = m_sub;
  args.acc_stages = (2 * m_sub * d->block_n <= 512) ? 2 : 1;
  args.act = d->act;
  args.stages = d->stages;
  args.bias = d->bias;
  args.out = d->out;
  args.stat.partial = d->stat_partial;
  args.stat.rows_per_img = d->stat_rows_per_img;
  args.stat.row0 = d->stat_row0;
  args.inbwd.mode = d->inbwd_mode;
  args.inbwd.has_add = d->inbwd_has_add;
  args.inbwd.h = d->inbwd_h;
  args.inbwd.w = d->inbwd_w;
  args.inbwd.halo = d->inbwd_halo;
  args.inbwd.add_shift = d->inbwd_add_halo - d->inbwd_halo;
  args.inbwd.n_tensors = 1 + (d->inbwd_mode == 2 ? 1 : 0) + (d->inbwd_has_add ? 1 : 0);
  {
    const char* dbg = getenv("FPG_INBWD_DEBUG");
    args.inbwd.debug = dbg ? atoi(dbg) : 0;
  }
  const bool inbwd = d->inbwd_mode != 0;
  CUtensorMap emap[6];  // z, z1, zprev, zprev1, add, add1 (copies of the activation map when unused)
  for (int i = 0; i < 6; ++i) emap[i] = amap;
  if (inbwd) {
    FPG_REQUIRE((d->inbwd_mode == 1 || d->inbwd_mode == 2) && !d->cta_pair && d->stat_partial != nullptr &&
                    d->out.fp32 == FPG_DT_BF16 && d->out.mul_y == 1 && d->out.mul_x == 1 && d->cblk == 64 &&
                    d->block_n % 64 == 0 && d->bias == nullptr && d->act == FPG_ACT_NONE,
                "InstanceNorm-backward reductions: 1-CTA stride-1 bf16 launch, block_n a multiple of 64, with the "
                "statistics buffer");
    const fpg_tmap* src[6] = {&d->inbwd_z,    &d->inbwd_z1,   &d->inbwd_prev,
                              &d->inbwd_prev1, &d->inbwd_add, &d->inbwd_add1};
    for (int i = 0; i < 6; ++i) {
      const bool region1 = (i & 1) != 0;
      const bool used = (i < 2) || (i < 4 && d->inbwd_mode == 2) || (i >= 4 && d->inbwd_has_add);
      if (!used || (region1 && !two_regions)) continue;
      FPG_REQUIRE(src[i]->box[0] == 32 && src[i]->swizzle_bytes == 64 &&
                      src[i]->box[1] == static_cast<uint32_t>(region1 ? d->tile_w1 : d->tile_w) &&
                      src[i]->box[3] == static_cast<uint32_t>(region1 ? d->tile_h1 : d->tile_h),
                  "InstanceNorm-backward operand view %d: box {32, tile_w, 1, tile_h, 1}, 64-byte swizzle", i);
      rc = encode_tmap(src[i], &emap[i]);
      if (rc) return rc;
    }
  }
  args.stat.c_total = d->block_n * d->n_blocks;
  for (int i = 0; i < FPG_MAX_TAPS; ++i) args.taps[i] = d->taps[i];

  const int total_tiles = d->n_img * d->tiles_y * d->tiles_x * d->n_blocks +
                          (two_regions ? d->n_img * d->tiles_y1 * d->tiles_x1 * d->n_blocks : 0);
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(FPG_ENOTSUP, "no CUDA device");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d->cta_pair) {
    FPG_REQUIRE(m_sub == 1 && d->block_n % 32 == 0 && d->cblk >= 32, "cta_pair needs 128-pixel tiles, block_n % 32");
    const int clusters = total_tiles < sms / 2 ? total_tiles : sms / 2;
    const size_t smem2 =
        static_cast<size_t>(d->stages) * (16384 + d->block_n * 64) + (2 * d->stages + 4) * 8 + 16 + 1024;
    FPG_REQUIRE(smem2 <= 227 * 1024, "shared memory %zu", smem2);
    if (d->cblk == 64) {
      FPG_CUDA_CHECK(cudaFuncSetAttribute(igemm_fprop2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem2)));
      igemm_fprop2_kernel<64><<<2 * clusters, kFpropThreads, smem2, st>>>(amap, bmap, args);
    } else {
      FPG_CUDA_CHECK(cudaFuncSetAttribute(igemm_fprop2_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem2)));
      igemm_fprop2_kernel<32><<<2 * clusters, kFpropThreads, smem2, st>>>(amap, bmap, args);
    }
    FPG_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  const int grid = total_tiles < sms ? total_tiles : sms;
  if (const char* cap = getenv("FPG_FPROP_STAGES_CAP")) {  // experiment: effect of the ring depth
    const int c = atoi(cap);
    if (c >= 2 && c < args.stages) args.stages = c;
  }
  const size_t smem = static_cast<size_t>(d->stages) * (16384 * m_sub + d->block_n * 128) +
                      (inbwd ? kEStages * args.inbwd.n_tensors * kESliceBytes : 0) +
                      (2 * d->stages + 4 + 2 * kEStages) * 8 + 16 + 1024;
  FPG_REQUIRE(smem <= 227 * 1024, "shared memory %zu", smem);
#define FPG_LAUNCH_FPROP(CB, IB, THREADS)                                                                          \
  do {                                                                                                             \
    FPG_CUDA_CHECK(cudaFuncSetAttribute(igemm_fprop_kernel<CB, IB>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                        static_cast<int>(smem)));                                                  \
    FPG_CUDA_CHECK(launch_persistent(igemm_fprop_kernel<CB, IB>, dim3(grid), dim3(THREADS), smem, st, amap, bmap,  \
                                     amap1, emap[0], emap[1], emap[2], emap[3], emap[4], emap[5], args));          \
  } while (0)
  if (inbwd) {
    FPG_LAUNCH_FPROP(64, true, kFpropThreadsInBwd);
  } else if (d->cblk == 64) {
    FPG_LAUNCH_FPROP(64, false, kFpropThreads);
  } else if (d->cblk == 32) {
    FPG_LAUNCH_FPROP(32, false, kFpropThreads);
  } else {
    FPG_LAUNCH_FPROP(16, false, kFpropThreads);
  }
#undef FPG_LAUNCH_FPROP
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// The four parity-class launches of a stride-2 data gradient as ONE launch of igemm_s2cls_kernel. Returns 1 when the
// class plans do not qualify (the caller then launches them one by one): 1-CTA plans with one 64-channel chunk per K
// step, N = c_in <= 64 in one block, the same tiling for every class, no bias / activation, and all tap weights
// (taps x c_out x N bf16) fitting in shared memory beside at least two activation stages.
extern "C" int fpg_igemm_s2cls_launch(const fpg_igemm_fprop_desc* d, int32_t n_cls, void* stream) {
  FPG_REQUIRE(d != nullptr, "null descriptor");
  static const bool disabled = getenv("FPG_DISABLE_S2CLS") != nullptr;
  if (disabled || n_cls != 4) return 1;
  int n_slots = 0;
  for (int q = 0; q < 4; ++q) {
    const fpg_igemm_fprop_desc& c = d[q];
    if (c.cblk != 64 || c.cta_pair || c.n_blocks != 1 || c.block_n > 64 || c.block_n % 16 != 0 ||
        c.tile_h * c.tile_w != 128 || c.tiles_x1 > 0 || c.bias != nullptr || c.act != FPG_ACT_NONE ||
        c.inbwd_mode != 0 || c.c_per_tap % 64 != 0 || c.num_taps < 1 || c.num_taps > 4 ||
        c.out.mul_y != 2 || c.out.mul_x != 2)
      return 1;
    if (c.tile_h != d[0].tile_h || c.tile_w != d[0].tile_w || c.tiles_x != d[0].tiles_x ||
        c.tiles_y != d[0].tiles_y || c.block_n != d[0].block_n || c.c_per_tap != d[0].c_per_tap ||
        c.n_img != d[0].n_img || c.out.base != d[0].out.base || c.a.base != d[0].a.base ||
        c.stat_partial != d[0].stat_partial || c.out.fp32 != d[0].out.fp32)
      return 1;
    n_slots += c.num_taps * (c.c_per_tap / 64);
  }
  const int BN = d[0].block_n;
  const size_t resident = (static_cast<size_t>(n_slots) * BN * 128 + 1023) & ~size_t(1023);
  const size_t budget = 227 * 1024 - 1024 - 256 - kS2MaxOps * 12;
  if (resident + 2 * 16384 > budget) return 1;
  int stages = static_cast<int>((budget - resident) / 16384);
  if (stages > 6) stages = 6;

  S2Args args;
  memset(&args, 0, sizeof(args));
  args.n_img = d[0].n_img;
  args.tiles_y = d[0].tiles_y;
  args.tiles_x = d[0].tiles_x;
  args.tile_w_log2 = log2_exact(d[0].tile_w);
  args.tile_h = d[0].tile_h;
  args.tile_w = d[0].tile_w;
  FPG_REQUIRE(args.tile_w_log2 >= 0, "tile width %d", d[0].tile_w);
  args.block_n = BN;
  args.kchunks = d[0].c_per_tap / 64;
  args.stages = stages;
  static const bool row_store = getenv("FPG_S2CLS_ROWSTORE") != nullptr;  // A/B switch: row-per-thread stores
  args.line_store = row_store ? 0 : 1;
  int slot = 0;
  for (int q = 0; q < 4; ++q) {
    args.cls_taps[q] = d[q].num_taps;
    args.cls_slot0[q] = slot;
    for (int t = 0; t < d[q].num_taps; ++t) {
      const fpg_tap& tp = d[q].taps[t];
      FPG_REQUIRE(tp.c0 == 0 && tp.plane == 0, "class taps must address the stride-1 view of dy");
      int si = 0;
      while (si < args.n_shifts && (args.shift[si].ty != tp.dy || args.shift[si].tx != tp.dx)) ++si;
      if (si == args.n_shifts) {
        if (args.n_shifts == kS2MaxShifts) return 1;
        args.shift[si].ty = static_cast<int16_t>(tp.dy);
        args.shift[si].tx = static_cast<int16_t>(tp.dx);
        ++args.n_shifts;
      }
      auto& sh = args.shift[si];
      if (sh.n_users == 4) return 1;
      sh.u[sh.n_users].cls = static_cast<int16_t>(q);
      sh.u[sh.n_users].slot = static_cast<int16_t>(slot + t * args.kchunks);
      ++sh.n_users;
    }
    slot += d[q].num_taps * args.kchunks;
    args.off_y[q] = d[q].out.off_y;
    args.off_x[q] = d[q].out.off_x;
    args.stat_row0[q] = d[q].stat_row0;
  }
  args.out = d[0].out;
  args.stat.partial = d[0].stat_partial;
  args.stat.rows_per_img = d[0].stat_rows_per_img;
  args.stat.row0 = 0;
  args.stat.c_total = BN;
  CUtensorMap amap, bmap[4];
  int rc = encode_tmap(&d[0].a, &amap);
  if (rc) return rc;
  for (int q = 0; q < 4; ++q) {
    rc = encode_tmap(&d[q].b, &bmap[q]);
    if (rc) return rc;
  }
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(FPG_ENOTSUP, "no CUDA device");
  const int total_tiles = args.n_img * args.tiles_y * args.tiles_x;
  const int grid = total_tiles < sms ? total_tiles : sms;
  if (n_slots > kS2MaxOps) return 1;
  const size_t smem = resident + static_cast<size_t>(stages) * 16384 + (2 * stages + 5) * 8 + kS2MaxOps * 12 + 16 + 1024;
  FPG_CUDA_CHECK(cudaFuncSetAttribute(igemm_s2cls_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
  FPG_CUDA_CHECK(launch_persistent(igemm_s2cls_kernel, dim3(grid), dim3(kFpropThreads), smem,
                                   static_cast<cudaStream_t>(stream), amap, bmap[0], bmap[1], bmap[2], bmap[3], args));
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

static int launch_wgrad_pair(const fpg_igemm_wgrad_desc* d, void* stream) {
  FPG_REQUIRE(d->x_ca == 64 && d->y_ca == 64 && d->x_atoms == 4 && d->y_atoms == 4 && d->x_groups == 1 &&
                  d->y_groups == 1 && d->y_sets == 2 && !d->x_taps_mode && !d->y_taps_mode && d->x_is_dy &&
                  !d->x_shift_atoms && !d->y_shift_atoms && d->y_shifts <= 1,
              "cta_pair wgrad: 256 x 256 channels, two taps per item");
  FPG_REQUIRE(d->tile_h * d->tile_w == 64 && d->ws != nullptr && d->splits >= 1 && d->stages >= 2 && d->stages <= 4 &&
                  (!(d->y_ntaps & 1) || (d->last_splits >= 1 && d->last_splits <= d->splits)),
              "cta_pair wgrad geometry");
  CUtensorMap xmap, ymap;
  int rc = encode_tmap(&d->x, &xmap);
  if (rc) return rc;
  rc = encode_tmap(&d->y, &ymap);
  if (rc) return rc;
  Wgrad2Args args;
  args.n_img = d->n_img;
  args.kt_y = d->kt_y;
  args.kt_x = d->kt_x;
  args.tile_h = d->tile_h;
  args.tile_w = d->tile_w;
  args.splits = d->splits;
  args.last_splits = d->last_splits;
  args.stages = d->stages;
  args.ntaps = d->y_ntaps;
  args.items = (d->y_ntaps + 1) / 2;
  const int clusters = (d->y_ntaps / 2) * d->splits + (d->y_ntaps & 1) * args.last_splits;
  args.ws = d->ws;
  for (int i = 0; i < FPG_MAX_TAPS; ++i) args.y_taps[i] = d->y_taps[i];
  const size_t smem = static_cast<size_t>(d->stages) * 6 * 8192 + (2 * 8 + 2) * 8 + 16 + 1024;
  FPG_CUDA_CHECK(cudaFuncSetAttribute(igemm_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
  igemm_wgrad2_kernel<<<2 * clusters, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(xmap, ymap,
                                                                                                       args);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int fpg_igemm_wgrad_launch(const fpg_igemm_wgrad_desc* d, void* stream) {
  FPG_REQUIRE(d != nullptr, "null descriptor");
  if (d->cta_pair) return launch_wgrad_pair(d, stream);
  const int M = d->x_atoms * d->x_ca, N = d->y_atoms * d->y_ca;
  const int ysh = d->y_shifts > 1 ? d->y_shifts : 1, ysets = d->y_sets > 1 ? d->y_sets : 1;
  const int ys = ysh * ysets;
  FPG_REQUIRE(M == 64 || M == 128 || (M == 256 && d->x_ca == 64 && !d->x_shift_atoms), "wgrad M %d", M);
  FPG_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "wgrad N %d", N);
  FPG_REQUIRE((M > 128 ? 2 : 1) * ys * N <= 512, "accumulators do not fit TMEM");
  FPG_REQUIRE(d->tile_h * d->tile_w == 64, "k tile %dx%d", d->tile_h, d->tile_w);
  const bool shifted = d->x_shift_atoms || d->y_shift_atoms || ys > 1;
  FPG_REQUIRE(!shifted || d->tile_h == 1, "shifted operands need one-row k tiles");
  FPG_REQUIRE(!(d->y_shift_atoms && ysh > 1) && (ysets == 1 || d->y_shift_atoms),
              "y_shift_atoms excludes y_shifts; y_sets needs y_shift_atoms");
  FPG_REQUIRE(d->stages >= 2 && d->stages <= 8 && d->splits >= 1, "stages %d splits %d", d->stages, d->splits);
  FPG_REQUIRE(d->ws != nullptr, "null workspace");
  CUtensorMap xmap, ymap;
  int rc = encode_tmap(&d->x, &xmap);
  if (rc) return rc;
  rc = encode_tmap(&d->y, &ymap);
  if (rc) return rc;
  WgradArgs args;
  args.x_ca = d->x_ca;
  args.y_ca = d->y_ca;
  args.x_atoms = d->x_atoms;
  args.y_atoms = d->y_atoms;
  args.x_groups = d->x_groups;
  args.y_groups = d->y_groups;
  args.x_taps_mode = d->x_taps_mode;
  args.y_taps_mode = d->y_taps_mode;
  args.x_ntaps = d->x_ntaps;
  args.y_ntaps = d->y_ntaps;
  args.n_img = d->n_img;
  args.kt_y = d->kt_y;
  args.kt_x = d->kt_x;
  args.tile_h = d->tile_h;
  args.tile_w = d->tile_w;
  args.splits = d->splits;
  args.stages = d->stages;
  args.x_shift_atoms = d->x_shift_atoms ? 1 : 0;
  args.y_shift_atoms = d->y_shift_atoms ? 1 : 0;
  args.y_shifts = ysh;
  args.y_sets = ysets;
  // box extents in pixels: 64 (+ shift range)
  const uint32_t x_px = 64u + (d->x_shift_atoms ? d->x_atoms - 1 : 0);
  const uint32_t y_px = 64u + (d->y_shift_atoms ? d->y_atoms - 1 : 0) + (ysh - 1);
  FPG_REQUIRE(d->x.box[1] * d->x.box[3] == x_px && d->y.box[1] * d->y.box[3] == y_px, "box extents %u %u",
              d->x.box[1], d->y.box[1]);
  args.x_box_bytes = x_px * d->x_ca * 2u;
  args.y_box_bytes = y_px * d->y_ca * 2u;
  args.x_atom_stride = (args.x_box_bytes + 1023u) & ~1023u;
  args.y_atom_stride = (args.y_box_bytes + 1023u) & ~1023u;
  args.x_stage_bytes = args.x_atom_stride * (d->x_shift_atoms ? 1 : d->x_atoms);
  args.y_stage_bytes = args.y_atom_stride * (d->y_shift_atoms ? ysets : d->y_atoms);
  args.ws = d->ws;
  for (int i = 0; i < FPG_MAX_TAPS; ++i) {
    args.x_taps[i] = d->x_taps[i];
    args.y_taps[i] = d->y_taps[i];
  }
  const int NX = d->x_taps_mode ? d->x_groups : d->x_groups * d->x_ntaps;
  const int NY = d->y_taps_mode ? (d->y_groups + ysets - 1) / ysets : d->y_groups * d->y_ntaps;
  const int grid = NX * NY * d->splits;
  const size_t smem = static_cast<size_t>(d->stages) * (args.x_stage_bytes + args.y_stage_bytes) +
                      (2 * d->stages + 2) * 8 + 16 + 1024;
  FPG_REQUIRE(smem <= 227 * 1024, "shared memory %zu", smem);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FPG_CUDA_CHECK(cudaFuncSetAttribute(igemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
  FPG_CUDA_CHECK(launch_persistent(igemm_wgrad_kernel, dim3(grid), dim3(kThreads), smem, st, xmap, ymap, args));
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}
