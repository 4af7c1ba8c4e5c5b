/*
 * fpg.h -- C ABI of libfpg_b200.so, the sm_100a kernel library behind the GAN training-step hot path of
 * Natasha-R/Flood-Prediction-GAN (reference: models/model_architectures.py, models/model.py:598-758).
 *
 * The reference has no FFI of its own: every operator below replaces a torch.nn / ATen call site on that path
 * (cited per function as file:line relative to the reference root). Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - activations are NHWC bf16 ("pixel-major": the channel run of one pixel is contiguous), channel counts
 *     padded to a multiple of 16 with the padding kept at exactly 0;
 *   - buffers are owned by the caller; kernels borrow them for the launch, allocate nothing, keep no state;
 *   - `stream` is a cudaStream_t passed as void*; all launches are asynchronous on it;
 *   - return value: 0 on success, otherwise a cudaError_t / CUresult code or a negative FPG_E* code.
 *     fpg_last_error() returns a human-readable message for the calling thread. No exceptions cross the ABI.
 *   - there is no CPU fallback: on a machine without an sm_100 GPU every compute call fails.
 */
#ifndef FPG_H_
#define FPG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPG_ABI_VERSION 2
#define FPG_MAX_TAPS 64

#define FPG_EINVAL (-22)
#define FPG_ENOTSUP (-95)

/* activation codes for fused epilogues / elementwise passes */
#define FPG_ACT_NONE 0
#define FPG_ACT_RELU 1
#define FPG_ACT_LEAKY 2 /* LeakyReLU(0.2) */
#define FPG_ACT_TANH 3

int fpg_abi_version(void);
const char* fpg_last_error(void);
/* number of SMs of the current device (148 on B200); <0 on error */
int fpg_sm_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Generic implicit-GEMM descriptors. The semantic conv entry points below are thin planners that fill these and
 * launch; the *_plan variants only fill the descriptor (pure host code, usable without a GPU) so that tests can
 * check the gather geometry against an independent CPU interpreter.
 * ---------------------------------------------------------------------------------------------------------- */

/* One filter tap as TMA coordinate offsets into a 5-D view (c, x, plane, y, n) of an NHWC tensor. */
typedef struct {
  int32_t c0;    /* offset on dim 0 (channel, or x-parity * channel-stride for stride-2 views) */
  int32_t dx;    /* offset on dim 1 (x, or x/2) */
  int32_t plane; /* coordinate on dim 2 (row-parity plane; 0 for stride-1 views) */
  int32_t dy;    /* offset on dim 3 (y, or y/2) */
} fpg_tap;

/* A bf16 tensor view for TMA: rank <= 5, dim 0 contiguous. strides[i] is the byte stride of dim i+1. */
typedef struct {
  void* base;
  int32_t rank;
  int32_t swizzle_bytes; /* 32 / 64 / 128 == box[0] * 2 */
  uint64_t dims[5];
  uint64_t strides[4];
  uint32_t box[5];
} fpg_tmap;

/* out[n, y*mul_y+off_y, x*mul_x+off_x, k] for tile pixel (n, y, x), element strides */
typedef struct {
  void* base;
  int64_t stride_n, stride_y, stride_x; /* in elements */
  int32_t mul_y, off_y, mul_x, off_x;
  int32_t valid_h, valid_w; /* tile pixels with y >= valid_h or x >= valid_w are not stored */
  int32_t fp32;             /* element type of the output: FPG_DT_BF16 / FPG_DT_FP32 / FPG_DT_FP16 */
} fpg_out_view;

/* D[pixel, k] = sum_{tap, c} A[pixel + tap, c] * B[k, tap*C + c]  (+bias, activation) */
typedef struct {
  fpg_tmap a;      /* gathered operand: box = {cblk, tile_w, 1, tile_h, 1}, tile_w*tile_h == 128 or 256 */
  fpg_tmap b;      /* weight matrix [n_total rows][num_sub*cblk], K-major: box = {cblk, block_n} */
  int32_t cblk;    /* channels per TMA sub-load: 16 / 32 / 64 */
  int32_t c_per_tap; /* channels per tap (multiple of cblk) */
  int32_t num_taps;  /* taps listed in taps[] (incl. zero-weight padding taps) */
  int32_t num_sub;   /* num_taps * c_per_tap / cblk; multiple of 64/cblk */
  int32_t block_n;   /* N tile: multiple of 16, <= 256 */
  int32_t n_blocks;  /* n_total / block_n */
  int32_t n_img, tiles_y, tiles_x, tile_h, tile_w;
  int32_t act;
  int32_t stages;
  int32_t cta_pair;  /* 1: 2-CTA kernel (tcgen05 cta_group::2): a tile is 2*tile_h x tile_w pixels, the CTA of cluster
                        rank r takes rows [r*tile_h, (r+1)*tile_h) and loads weight rows [r*block_n/2, ...): b.box[1] ==
                        block_n / 2; tiles_y counts pair tiles */
  const float* bias; /* [n_total] or NULL */
  fpg_out_view out;
  fpg_tap taps[FPG_MAX_TAPS];
  /* Optional second tile region (tiles_x1 > 0): the columns x >= x_org1 are covered by tiles of tile_w1 x tile_h1
   * pixels gathered through `a1` (same view, different box), so that an extent just above a multiple of 64 (66 = a
   * 64-pixel row with its reflect halo) does not force a tile shape that wastes a third of every tile; region 0 then
   * covers x < x_org1 only (tiles_x * tile_w == x_org1). Not combined with cta_pair. */
  fpg_tmap a1;
  int32_t tiles_y1, tiles_x1, tile_h1, tile_w1, x_org1;
  /* Optional normalisation statistics from the epilogue (stat_partial != NULL): every epilogue warp writes the
   * {sum, sum of squares} over its 32 pixel rows of the bf16-rounded stored values of every column to
   * stat_partial[image][stat_row0 + local row][column][2]; local row = (tile index inside the image) * 4 + warp.
   * fpg_conv_stats_rows() gives the rows one launch contributes per image. */
  float* stat_partial;
  int32_t stat_rows_per_img, stat_row0;
  /* Optional InstanceNorm-BACKWARD reductions (inbwd_mode != 0, needs stat_partial): the launch produces dz, the
   * gradient w.r.t. the (reflect-haloed) input z of a stride-1 convolution whose input is the output of an
   * InstanceNorm block. Instead of {sum, sum of squares} the epilogue accumulates per column what that norm's backward
   * needs, from z itself (the tensor saved for the weight gradient; same geometry as dz):
   *   inbwd_mode 1, z = relu(zhat):                  {sum f [z > 0], sum f z}    = {sum g', sum g' zhat}
   *   inbwd_mode 2, z = zprev + zhat (residual add): {sum f, sum f (z - zprev)}  = {sum g,  sum g zhat}
   * f = the stored value. The sums are linear in dz and z's halo mirrors its interior, so the halo needs no fold first.
   * inbwd_has_add: the skip-connection gradient (view inbwd_add) is added to the INTERIOR outputs before they are
   * stored and reduced; it is read at coordinates shifted by inbwd_add_halo - inbwd_halo.
   * The views are 5-D like `a` with boxes {32, tile_w, 1, tile_h, 1} (…1: tile shape of the second region). */
  int32_t inbwd_mode, inbwd_has_add;
  int32_t inbwd_h, inbwd_w, inbwd_halo, inbwd_add_halo;
  fpg_tmap inbwd_z, inbwd_z1, inbwd_prev, inbwd_prev1, inbwd_add, inbwd_add1;
} fpg_igemm_fprop_desc;

/* D_item[m, n] = sum_{pixel} X[pixel + xtap, xc + m] * Y[pixel + ytap, yc + n], split over pixel ranges,
 * partial tiles written as fp32 to ws[split][item][M][y_shifts * N].
 * Shifted operands (tile_h == 1): because shared-memory descriptors may start at any pixel row of a TMA-written box,
 * several filter taps that differ by whole pixels in x can be served by ONE box of tile_w + (taps - 1) pixels:
 *   x_shift_atoms / y_shift_atoms = 1: the atoms of the tile are the pixel shifts 0 .. atoms-1 of one box (one MMA,
 *       descriptor atom stride = one pixel row); the tap tables must list dx(first) + a for atom a;
 *   y_shifts = g > 1: g MMA groups per Y tile, group j reads the Y tile j pixels later and accumulates into its own
 *       columns [j*N, (j+1)*N): one box per channel atom instead of g;
 *   y_sets = g > 1 (with y_shift_atoms): g shift-atom boxes per stage, i.e. an item covers the g consecutive tap
 *       groups yi*g .. yi*g + g-1 of the Y table, each with its own accumulator columns: the X tile is loaded once
 *       for all of them (the X stream otherwise repeats per tap group and saturates L2 -> SM bandwidth). */
typedef struct {
  fpg_tmap x, y;          /* box = {ca, tile_w (+ shift extent), 1, tile_h, 1}, tile_w*tile_h == 64 */
  int32_t x_ca, y_ca;     /* atom width in channels: 16 / 32 / 64 */
  int32_t x_atoms, y_atoms; /* M = x_atoms*x_ca in {64,128,256}; N = y_atoms*y_ca, multiple of 16, <= 256 */
  int32_t x_groups, y_groups;
  int32_t x_taps_mode, y_taps_mode; /* 1: atom index enumerates taps; 0: atom index enumerates channel chunks */
  int32_t x_ntaps, y_ntaps;
  int32_t n_img, kt_y, kt_x, tile_h, tile_w;
  int32_t splits, stages;
  int32_t x_is_dy; /* 1: X operand is the output gradient (rows of D are output channels); 0: X is the input */
  int32_t tap_on_x; /* informational: 1 if the filter taps are enumerated by the X operand only */
  int32_t taps_r, taps_s; /* filter size: tap ids >= taps_r*taps_s (or < 0) are padding and are dropped */
  int32_t x_shift_atoms, y_shift_atoms, y_shifts, y_sets;
  int32_t cta_pair; /* 1: 2-CTA kernel for 256 x 256-channel layers: an item is a pair of taps (y_sets == 2, the sets
                       enumerate taps, the 4 atoms channels), CTA r of the pair loads its half of both operands */
  int32_t last_splits; /* cta_pair with an odd tap count: k-range splits of the last, single-tap item (<= splits) */
  float* ws;
  fpg_tap x_taps[FPG_MAX_TAPS];
  fpg_tap y_taps[FPG_MAX_TAPS];
  /* filter tap id r*S+s of element (m, n) = x_tap_rs[x tap index] + y_tap_rs[y tap index] + shift group */
  int16_t x_tap_rs[FPG_MAX_TAPS];
  int16_t y_tap_rs[FPG_MAX_TAPS];
} fpg_igemm_wgrad_desc;

/* Row-stationary implicit GEMM for stride-1 R x S filters on wide images (7x7 stem / heads): a CTA tile is
 * tile_rows output rows x 128 pixels with one TMEM accumulator per output row. Every input row of the halo patch is
 * loaded ONCE as a (128 + S - 1)-pixel TMA box and feeds up to tile_rows x S MMAs whose shared-memory descriptors start
 * at shifted pixel rows of that box (the 128B/64B/32B swizzles are functions of the address, so any pixel-row start is
 * legal); every filter row of B is loaded once per tile and reused by all accumulators.
 *   D[(y, x), k] = sum_{i < rows, j < cols, c} A[y + dy0 + i, x + dx0 + j, c] * B[k, tap_of[i*cols + j]*cblk + c]  */
typedef struct {
  fpg_tmap a;          /* 5-D stride-1 activation view, box = {cblk, 128 + cols - 1, 1, 1, 1} */
  fpg_tmap b;          /* weight matrix [block_n rows][num_taps_padded * cblk], K-major: box = {cblk, block_n} */
  int32_t cblk;        /* channels per tap == channels of A: 16 / 32 / 64 */
  int32_t block_n;     /* N: multiple of 16, <= 128 */
  int32_t rows, cols;  /* distinct row / column offsets of the filter (R, S) */
  int32_t dy0, dx0;    /* smallest row / column offset */
  int32_t tile_rows;   /* accumulators per tile: 2 * tile_rows * block_n <= 512 */
  int32_t n_img, tiles_y, tiles_x;
  int32_t act;
  int32_t a_stages;    /* ring of input-row slots */
  int32_t b_stages;    /* ring of filter-row slots; == rows: the whole filter stays resident for the CTA's lifetime */
  const float* bias;
  fpg_out_view out;
  int16_t tap_of[FPG_MAX_TAPS]; /* packed tap index of grid position (i, j) */
} fpg_igemm_rows_desc;

int fpg_igemm_fprop_launch(const fpg_igemm_fprop_desc* d, void* stream);
int fpg_igemm_wgrad_launch(const fpg_igemm_wgrad_desc* d, void* stream);
int fpg_igemm_rows_launch(const fpg_igemm_rows_desc* d, void* stream);
/* The n_cls == 4 output-parity class plans of a stride-2 data gradient / transposed-convolution forward
 * (fpg_conv2d_dgrad_plan) as ONE launch: every tap's weights resident in shared memory, each shifted activation tile
 * loaded once for all classes. Returns 0 (launched), 1 (plans do not qualify: launch them one by one) or an error. */
int fpg_igemm_s2cls_launch(const fpg_igemm_fprop_desc* descs, int32_t n_cls, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Convolution family (replaces nn.Conv2d / nn.ConvTranspose2d forward and aten::convolution_backward;
 * model_architectures.py:312-334 (generator), :407-416 (residual blocks), :424-437 (PatchGAN)).
 * ---------------------------------------------------------------------------------------------------------- */

#define FPG_DT_BF16 0
#define FPG_DT_FP32 1
#define FPG_DT_FP16 2

/* An NHWC activation buffer (bf16 unless `fp32` says otherwise). `halo` pixels of materialised border surround the h x w interior
 * (reflect halo written by the producer); c_stride >= c is the per-pixel element stride, `data` points at
 * channel 0 of the first halo pixel. */
typedef struct {
  void* data;
  int32_t n, h, w, c;
  int32_t c_stride;
  int32_t halo;
  int32_t fp32; /* element type FPG_DT_*: 0 = bf16 (tensor-core operands, gradients), 1 = fp32 (network heads),
                   2 = fp16 (pre-normalisation conv outputs and the residual skip stream: tensors that only elementwise
                   kernels read keep 3 more mantissa bits at the same 2 bytes/element; stores saturate) */
} fpg_act;

typedef struct {
  int32_t r, s;     /* filter size */
  int32_t stride;   /* 1 or 2 */
  int32_t pad;      /* zero padding (TMA out-of-bounds fill); reflect padding is a materialised input halo */
  int32_t c_in;     /* padded input channels of the packed weight */
  int32_t c_out;    /* padded output channels of the packed weight */
} fpg_conv_geom;

/* y = act(conv(x, w) + bias). w_packed: bf16 [c_out][r*s (padded to whole K stages)][c_in], see fpg_pack_weights.
 * Reads x at interior + halo (x.halo must equal the reflect pad of the layer, or 0).
 * nn.Conv2d call sites: model_architectures.py:343-345,353,369,414,416,424-437 */
int fpg_conv2d_fprop(const fpg_act* x, const void* w_packed, const float* bias, int act, const fpg_conv_geom* g,
                     const fpg_act* y, void* stream);
int fpg_conv2d_fprop_plan(const fpg_act* x, const void* w_packed, const float* bias, int act,
                          const fpg_conv_geom* g, const fpg_act* y, int sm_count, fpg_igemm_fprop_desc* out_desc);

/* dx = conv_backward_data(dy, w). g describes the FORWARD conv (y = conv(x)); w_packed_t is the dgrad packing
 * bf16 [c_in][taps][c_out] produced by fpg_pack_weights_dgrad. stride 1: one launch; stride 2: one launch per
 * output-parity class (descs[0..3]). dx covers interior + halo of the forward input when dx->halo > 0
 * (gradient w.r.t. the reflect-padded tensor; the halo fold happens in fpg_instnorm_bwd / fpg_halo_fold).
 * Also the forward of nn.ConvTranspose2d (model_architectures.py:350-351,367-368) with x := dy.
 * aten::convolution_backward (input grad): model.py:632,645 */
int fpg_conv2d_dgrad(const fpg_act* dy, const void* w_packed_t, const float* bias, int act, const fpg_conv_geom* g,
                     const fpg_act* dx, void* stream);
int fpg_conv2d_dgrad_plan(const fpg_act* dy, const void* w_packed_t, const float* bias, int act,
                          const fpg_conv_geom* g, const fpg_act* dx, int sm_count, fpg_igemm_fprop_desc* out_descs,
                          int* n_descs);
/* kernels fpg_conv2d_dgrad launches for this layer (1, or 4 when the parity classes run one by one) */
int32_t fpg_conv2d_dgrad_launches(const fpg_act* dy, const fpg_conv_geom* g, const fpg_act* dx);

/* Row-stationary plan of the same operation as fpg_conv2d_fprop (dgrad == 0: a = x, out = y) or fpg_conv2d_dgrad
 * (dgrad != 0: a = dy, out = dx, w_packed = the dgrad packing). Returns 0 and fills *out_desc when that path applies
 * (stride 1, filter larger than 1x1, one channel chunk of 16/32/64 on the gathered side, at most 128 channels on the
 * produced side, at least 96 output columns), 1 when it does not (fpg_conv2d_fprop / _dgrad then use the tiled
 * kernel), another code on invalid arguments. fpg_conv2d_fprop and fpg_conv2d_dgrad call this first. */
int fpg_conv2d_rows_plan(const fpg_act* a, const void* w_packed, const float* bias, int act, const fpg_conv_geom* g,
                         const fpg_act* out, int dgrad, int sm_count, fpg_igemm_rows_desc* out_desc);

/* Convolution + normalisation statistics in one pass: as fpg_conv2d_fprop / fpg_conv2d_dgrad, and the epilogue also
 * writes per-warp partial column sums of the stored output to stat_partial (>= 1.02 * n * rows * c_out_padded * 2
 * + 4096 floats: the tail is scratch of the finalisation; rows = fpg_conv_stats_rows(); 0 rows: this layer runs on the
 * row-stationary kernel, use fpg_instnorm_stats).
 * fpg_instnorm_stats_finalize reduces them to stats[(n*c + ch)*2] = {mean, rstd} (n = 1 and count = N*H*W: BatchNorm).
 * Replaces the separate statistics pass of nn.InstanceNorm2d / nn.BatchNorm2d (model_architectures.py:313-333). */
int32_t fpg_conv_stats_rows(const fpg_act* a, const fpg_conv_geom* g, const fpg_act* out, int dgrad);
int fpg_conv2d_fprop_stats(const fpg_act* x, const void* w_packed, const float* bias, int act, const fpg_conv_geom* g,
                           const fpg_act* y, float* stat_partial, void* stream);
int fpg_conv2d_dgrad_stats(const fpg_act* dy, const void* w_packed_t, const float* bias, int act,
                           const fpg_conv_geom* g, const fpg_act* dx, float* stat_partial, void* stream);
int fpg_instnorm_stats_finalize(const float* stat_partial, int32_t rows_per_img, int32_t n, int32_t c,
                                int64_t count_per_img, float eps, float* stats, void* stream);

/* Data gradient of a stride-1 convolution whose input z (with its reflect halo; the tensor saved for the weight
 * gradient) is the output of an InstanceNorm block, fused with the reduction pass of that norm's backward:
 * dx (incl. halo) = conv_backward_data(dy, w) [+ add on the interior], and stat_partial receives the per-warp partials
 * of the two plane sums (see fpg_igemm_fprop_desc.inbwd_*; sized as for fpg_conv2d_dgrad_stats):
 *   zprev == NULL: z = relu(IN(y))           (first convolution's output inside a residual block; the trunk input)
 *   zprev != NULL: z = zprev + IN(y)         (block output; zprev = the block's input, same geometry as z)
 * fpg_instnorm_bwd_sums_finalize turns them into red[(n*c + ch)*2] = {mean g', mean g' * zhat} over the h*w plane;
 * fpg_instnorm_bwd_apply then folds dx's halo in place and applies dy = rstd * (g' - mean g' - zhat * mean g' zhat):
 * together they equal fpg_instnorm_bwd(dz = dx, dz2 = add, ...) without its streamed reduction pass (the dominant
 * elementwise cost of the residual trunk's backward: model_architectures.py:408-418 under autograd).
 * Returns 1 when the layer does not plan as one tiled 1-CTA launch with block_n % 64 == 0 (caller falls back). */
int fpg_conv2d_dgrad_inbwd(const fpg_act* dy, const void* w_packed_t, const fpg_conv_geom* g, const fpg_act* dx,
                           const fpg_act* z, const fpg_act* zprev, const fpg_act* add, float* stat_partial,
                           int32_t* rows_per_img, void* stream);
int fpg_instnorm_bwd_sums_finalize(const float* stat_partial, int32_t rows_per_img, int32_t n, int32_t c,
                                   int64_t count_per_img, float* red, void* stream);

/* dw = conv_backward_weight(x, dy): fp32 gradient written (not accumulated) in the reference parameter layout.
 *   dw[ko*dw_stride_k + ci*dw_stride_c + (r*S+s)] for ko < k_valid, ci < c_valid.
 * nn.Conv2d weight [K][C][R][S]: dw_stride_k = C*R*S, dw_stride_c = R*S;
 * nn.ConvTranspose2d weight [Cin_T][Cout_T][R][S] (call with x := dy_T, dy := x_T): see INTEGRATION.md.
 * ws: fp32 workspace of at least fpg_conv2d_wgrad_ws_bytes() bytes.
 * aten::convolution_backward (weight grad): model.py:632,645 */
int fpg_conv2d_wgrad(const fpg_act* x, const fpg_act* dy, const fpg_conv_geom* g, float* dw, int64_t dw_stride_k,
                     int64_t dw_stride_c, int32_t k_valid, int32_t c_valid, float* ws, void* stream);
int fpg_conv2d_wgrad_plan(const fpg_act* x, const fpg_act* dy, const fpg_conv_geom* g, int sm_count,
                          fpg_igemm_wgrad_desc* out_desc);
int64_t fpg_conv2d_wgrad_ws_bytes(const fpg_act* x, const fpg_act* dy, const fpg_conv_geom* g, int sm_count);

/* Weight repack fp32 parameter -> bf16 GEMM operand.
 * fprop packing:  dst[k][t][c] = src[k*src_stride_k + c*src_stride_c + t]          (t = r*S+s)
 * dgrad packing:  dst[c][t'][k] with the taps flipped / parity-grouped as fpg_conv2d_dgrad expects.
 * Rows/taps/channels beyond the valid counts are written as 0. */
int fpg_pack_weights(const float* src, int64_t src_stride_k, int64_t src_stride_c, int32_t k_valid, int32_t c_valid,
                     const fpg_conv_geom* g, void* dst, void* stream);
int fpg_pack_weights_dgrad(const float* src, int64_t src_stride_k, int64_t src_stride_c, int32_t k_valid,
                           int32_t c_valid, const fpg_conv_geom* g, void* dst, void* stream);
/* Batched repack: every packed operand (and padded bias vector) of a network in ONE launch after the optimiser
 * step. The job table is built once on the host with fpg_pack_jobs (up to 5 jobs per layer: the fprop operand and one
 * dgrad operand per parity class; pass NULL for an operand that is not needed) / fpg_pack_job_copy_f32, copied to the
 * device together with a block table: block b packs elements [2048*block_first[b], +2048) of job block_job[b];
 * fpg_pack_job_blocks(job) blocks cover a job. */
typedef struct {
  const float* src;
  void* dst;
  int32_t rows, taps, cols; /* dst[rows][taps][cols] */
  int32_t rows_valid, cols_valid;
  int32_t dst_fp32;         /* 0: bf16 destination, 1: fp32 destination (bias vectors) */
  int64_t src_stride_row, src_stride_col;
  int8_t src_tap[FPG_MAX_TAPS]; /* source tap r*S+s of every destination tap, -1 = zero */
} fpg_pack_job;
int fpg_pack_jobs(const float* src, int64_t src_stride_k, int64_t src_stride_c, int32_t k_valid, int32_t c_valid,
                  const fpg_conv_geom* g, void* dst_fprop, void* dst_dgrad, fpg_pack_job* jobs /* >= 5 */,
                  int32_t* n_jobs);
int fpg_pack_job_copy_f32(const float* src, int32_t count_valid, float* dst, int32_t count_padded, fpg_pack_job* job);
int32_t fpg_pack_job_blocks(const fpg_pack_job* job);
int fpg_pack_weights_batched(const fpg_pack_job* jobs_dev, const int32_t* block_job_dev, const int32_t* block_first_dev,
                             int32_t n_blocks, void* stream);
/* Introspection of the dgrad packing: for parity class `cls` (0 for stride 1; (oy&1)*2+(ox&1) for stride 2) returns
 * the forward tap index r*S+s of every packed tap (-1 = zero padding tap), the padded tap count, the element offset
 * of the class matrix [c_in][taps][c_out] inside the packed buffer and the number of classes. */
int fpg_dgrad_class_info(const fpg_conv_geom* g, int cls, int32_t* src_tap /* [FPG_MAX_TAPS] */,
                         int32_t* num_taps_padded, int64_t* elem_offset, int32_t* num_classes);
/* bytes of the packed operands for geometry g */
int64_t fpg_packed_weight_bytes(const fpg_conv_geom* g);
int64_t fpg_packed_weight_dgrad_bytes(const fpg_conv_geom* g);

/* db[k] = sum over pixels of dy[.., k]  (bias gradient), k < k_valid; scratch: >= 592 * dy->c floats */
int fpg_bias_grad(const fpg_act* dy, float* db, int32_t k_valid, float* scratch, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * InstanceNorm + activation (+ residual, + reflect halo) -- nn.InstanceNorm2d(eps=1e-5, affine=False) followed by
 * F.relu / nn.LeakyReLU(0.2) / F.pad(reflect): model_architectures.py:313-333,342-352,408-416,430-436.
 * ---------------------------------------------------------------------------------------------------------- */

/* scratch floats needed by fpg_instnorm_stats / fpg_instnorm_bwd for activation y */
int64_t fpg_instnorm_scratch_floats(const fpg_act* y);
/* stats[(n*C + c)*2 + {0,1}] = {mean, rstd} over the h*w plane of y (biased variance, eps). Deterministic
 * two-stage reduction through `scratch`; `counters`: int32[4096] that is zero on entry and left zero (arrival counters
 * and release flags of the per-image rendezvous; shared by all instnorm calls of one stream; n <= 2048). */
int fpg_instnorm_stats(const fpg_act* y, float eps, float* stats, float* scratch, int32_t* counters, void* stream);
/* z = act((y - mean) * rstd) [+ residual]; written to z's interior and, if z->halo > 0, mirrored into its halo.
 * residual may be NULL; it is read at interior coordinates (its own halo is skipped). y and residual may be bf16 or
 * fp16 (FPG_DT_*). skip_out (may be NULL; halo-free, same geometry, bf16 or fp16) receives the same values before
 * they are rounded to z's bf16: the residual stream of the ResNet trunk (model_architectures.py:412-418) is carried
 * in fp16 beside the bf16 tensor-core operand, so that nine skip additions do not re-round it to 8 mantissa bits. */
int fpg_instnorm_apply(const fpg_act* y, const float* stats, int act, const fpg_act* residual, const fpg_act* z,
                       const fpg_act* skip_out, void* stream);
/* Backward of z = act(IN(y)). Upstream gradient g = fold(dz) + dz2, where dz is the gradient w.r.t. z INCLUDING
 * its halo when dz->halo > 0 (folded back onto the interior: backward of F.pad(reflect)) and dz2 (may be NULL,
 * halo ignored) is a second gradient branch (residual skip). Writes dy; if dres != NULL also writes g there
 * (gradient flowing on to the residual input). dz is CONSUMED: when dz->halo > 0 the mirror band of its interior may
 * be updated in place with the folded halo contributions. */
int fpg_instnorm_bwd(const fpg_act* dz, const fpg_act* dz2, const fpg_act* y, const float* stats, int act,
                     const fpg_act* dy, const fpg_act* dres, float* scratch, int32_t* counters, void* stream);
/* The apply half of fpg_instnorm_bwd with the reductions given: red[(n*c + ch)*2] = {mean g', mean g' * zhat}
 * (fpg_instnorm_bwd_sums_finalize). dz (with any second branch already merged) is CONSUMED: its halo is folded in place. */
int fpg_instnorm_bwd_apply(const fpg_act* dz, const fpg_act* y, const float* stats, const float* red, int act,
                           const fpg_act* dy, void* stream);
/* dx = fold(dz) * act'(z) for an activation without normalisation (PatchGAN model.0 LeakyReLU): z is the saved
 * activation output */
int fpg_act_bwd(const fpg_act* dz, const fpg_act* z, int act, const fpg_act* dx, void* stream);
/* c = fold(a) + b  (b may be NULL): backward of F.pad(reflect) plus a gradient accumulation */
int fpg_halo_fold(const fpg_act* a, const fpg_act* b, const fpg_act* c, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * BatchNorm2d in training mode + dropout -- the Pix2Pix U-Net and its BatchNorm PatchGAN
 * (model_architectures.py:9-85; nn.BatchNorm2d(eps=1e-5, momentum=0.1, affine), nn.Dropout(0.5)).
 * Batch statistics: fpg_instnorm_stats on the batch viewed as ONE image (n = 1, h = N*H) gives stats[c] = {mean, rstd}.
 * All tensors halo-free NHWC bf16; destinations / gradients may be channel slices (c_stride > c) of the U-Net's
 * concatenation buffers. In-place activations of the reference (:33-34,:63) give an encoder activation two
 * consumers, lrelu(e) (next down-conv) and relu(e) (skip connection): hence two outputs / two upstream gradients.
 * ---------------------------------------------------------------------------------------------------------- */
/* v = gamma * (y - mean) * rstd + beta (stats == NULL: v = y); v *= mask ? mask_scale : 0 (mask == NULL: none;
 * mask: uint8 [n*h*w][c]); z1 = act1(v); z2 = act2(v) if z2 != NULL */
int fpg_batchnorm_apply(const fpg_act* y, const float* stats, const float* gamma, const float* beta,
                        const uint8_t* mask, float mask_scale, int act1, const fpg_act* z1, int act2,
                        const fpg_act* z2, void* stream);
/* Backward of the above. Upstream g = (dz1 * act1'(v) + dz2 * act2'(v)) * mask (dz2 may be NULL).
 * stats != NULL: dy = gamma * rstd * (g - mean(g) - zhat * mean(g * zhat)), dbeta = sum g, dgamma = sum g * zhat
 * (written, or added if accumulate != 0; either may be NULL). stats == NULL: dy = g (y supplies the sign of v).
 * scratch: >= fpg_batchnorm_scratch_floats(y) floats. */
int fpg_batchnorm_bwd(const fpg_act* dz1, int act1, const fpg_act* dz2, int act2, const uint8_t* mask,
                      float mask_scale, const fpg_act* y, const float* stats, const float* gamma, const float* beta,
                      const fpg_act* dy, float* dgamma, float* dbeta, int accumulate, float* scratch, void* stream);
int64_t fpg_batchnorm_scratch_floats(const fpg_act* y);
/* running_mean = (1-m) running_mean + m mean; running_var = (1-m) running_var + m var * count/(count-1) */
int fpg_batchnorm_running_update(const float* stats, int32_t c, int64_t count, float eps, float momentum,
                                 float* running_mean, float* running_var, void* stream);
/* y = max over 2x2 windows, stride 2 (nn.MaxPool2d(2) of the segmentation U-Net, model_architectures.py:556) */
int fpg_maxpool2(const fpg_act* x, const fpg_act* y, void* stream);
/* mask[i] = 1 with probability keep (counter-based hash of seed and i: reproducible), else 0 */
int fpg_dropout_mask(uint8_t* mask, int64_t count, uint64_t seed, float keep, void* stream);
/* the same with seed = *seed_dev + seed_add read on the device (constant launch arguments: a captured training step
 * draws fresh masks on every replay) */
int fpg_dropout_mask_dev(uint8_t* mask, int64_t count, const uint64_t* seed_dev, uint64_t seed_add, float keep,
                         void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Input pipeline -- models/utils.py:19-67 (apply_transformations), models/data.py:57-78 (FloodDataset.__getitem__).
 *   fpg_resize_bicubic_aa: src_hwc = one decoded TIFF stack [in_h][in_w][in_c] fp32 (the layout tifffile.imread
 *     returns); keeps the `channels` (<= 16) input channels listed in channel_map (HOST array; the `topography`
 *     selection, utils.py:30-39), mirrors the columns when flip_w != 0 (np.fliplr, data.py:63-65), resamples with
 *     torchvision's Resize(BICUBIC, antialias=True) algorithm (utils.py:41-43; horizontal pass, then vertical, fp32)
 *     and writes dst_chw [channels][out_h][out_w] fp32. An axis whose size does not change is copied.
 *     scratch: fpg_resize_aa_scratch_bytes() bytes of device memory (256-byte aligned).
 *   fpg_tile_gather: images = DEVICE array of `batch` pointers to resident resized images [channels][height][width]
 *     fp32; sample b is window crop_index[b] (row-major, DEVICE int32 array) of a divisions x divisions grid of
 *     images[b] (utils.py:45-56), normalised as (v - mean) / stdv (utils.py:58-61);
 *     dst [batch][channels][height/divisions][width/divisions] fp32.
 * ---------------------------------------------------------------------------------------------------------- */
int64_t fpg_resize_aa_scratch_bytes(int32_t in_h, int32_t in_w, int32_t out_h, int32_t out_w, int32_t channels);
int fpg_resize_bicubic_aa(const float* src_hwc, int32_t in_h, int32_t in_w, int32_t in_c, const int32_t* channel_map,
                          int32_t channels, int32_t flip_w, int32_t out_h, int32_t out_w, float* dst_chw,
                          void* scratch, void* stream);
int fpg_tile_gather(const float* const* images, int32_t channels, int32_t height, int32_t width,
                    const int32_t* crop_index, int32_t batch, int32_t divisions, float mean, float stdv, float* dst,
                    void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Image-quality metrics of calculate_metrics -- models/model.py:367-371, 404-406 (torchmetrics 1.2.0, un-vendored:
 * PeakSignalNoiseRatio / StructuralSimilarityIndexMeasure / MultiScaleStructuralSimilarityIndexMeasure with
 * data_range=(0, 1)). Images are fp32 NCHW on the device.
 *   fpg_ssim_stats: out[img] = {mean SSIM, mean contrast sensitivity} over the image's C*(H-10)*(W-10) window
 *     positions (gaussian 11x11 window of `sigma`, constants (k1*data_range)^2 and (k2*data_range)^2, 5-pixel border
 *     cropped as torchmetrics does). scratch: fpg_ssim_scratch_bytes() bytes.
 *   fpg_avgpool2_f32: y[planes][h/2][w/2] = F.avg_pool2d(x, (2, 2)) -- the link between MS-SSIM scales.
 *   fpg_sq_err_sum: out[0] (double) = sum (clamp(a) - clamp(b))^2, fixed summation order -- the PSNR numerator.
 * ---------------------------------------------------------------------------------------------------------- */
int64_t fpg_ssim_scratch_bytes(int32_t n, int32_t c, int32_t h, int32_t w);
int fpg_ssim_stats(const float* pred, const float* target, int32_t n, int32_t c, int32_t h, int32_t w,
                   float data_range, float k1, float k2, float sigma, float* out, void* scratch, void* stream);
int fpg_avgpool2_f32(const float* x, float* y, int32_t planes, int32_t h, int32_t w, void* stream);
int64_t fpg_sq_err_scratch_bytes(void);
int fpg_sq_err_sum(const float* a, const float* b, int64_t count, float clamp_lo, float clamp_hi, double* out,
                   void* scratch, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Attention / content blend -- model_architectures.py:353-399.
 *   content: fp32, tanh already applied, 27 valid channels (9 RGB triplets) in a 32-channel buffer
 *   logits:  fp32, 10 valid channels in a 16-channel buffer (pre-softmax)
 *   image:   pre-flood RGB = channels 0..2 of `input` (interior of a possibly haloed buffer); input_lo_offset > 0
 *            (fp32 parity mode): plus channels input_lo_offset + 0..2, the low bf16 halves of the image
 *   out = sum_k content_k * a_k + image * a_10, written as bf16 into `out` channels [out_c0, out_c0+3)
 *   and as fp32 NCHW into out_nchw (may be NULL); mask_nhw (fp32 [n][h][w], may be NULL) gets a_10.
 * ---------------------------------------------------------------------------------------------------------- */
int fpg_blend_fwd(const fpg_act* content, const fpg_act* logits, const fpg_act* input, int32_t input_lo_offset,
                  const fpg_act* out, int32_t out_c0, float* out_nchw, float* mask_nhw, void* stream);
/* upstream gradient = dout_nchw (fp32 [n][3][h][w], may be NULL) + dout_nhwc channels [dout_c0, dout_c0+3)
 * (bf16, may be NULL) -> dcontent (pre-tanh gradient, bf16 32 ch), dlogits (bf16 16 ch) and, if not NULL,
 * dimage_nchw = gradient w.r.t. the pre-flood RGB (fp32 [n][3][h][w]; needed by the cycle models). */
int fpg_blend_bwd(const float* dout_nchw, const fpg_act* dout_nhwc, int32_t dout_c0, const fpg_act* content,
                  const fpg_act* logits, const fpg_act* input, const fpg_act* dcontent, const fpg_act* dlogits,
                  float* dimage_nchw, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Losses -- nn.MSELoss against torch.full(target) (model.py:626-631,641-642) and nn.L1Loss * weight (model.py:643).
 * Each writes the scalar loss (mean reduction, times `weight`) to *loss and the gradient of (grad_scale * loss).
 * ---------------------------------------------------------------------------------------------------------- */
/* logits: fp32 NHWC buffer with 1 valid channel (PatchGAN output). dlogits: bf16, same geometry (other channels 0).
 * scratch: >= logits->n floats; counter: one int32 that is zero on entry and left zero (ticket of the per-image CTAs:
 * the last one sums the per-image partials in image order). */
int fpg_mse_const_loss(const fpg_act* logits, float target, float weight, float grad_scale, float* loss,
                       const fpg_act* dlogits, float* scratch, int32_t* counter, void* stream);
/* pred/target: fp32 NCHW [count]; dpred (fp32, may be NULL) = grad_scale * weight * sign(pred-target) / count,
 * accumulated (+=) if accumulate != 0. per_image > 0: the target is the leading per_image elements of every
 * target_image_stride elements (real_image[:, :3] of a wider NCHW tensor: the cycle / identity losses,
 * model.py:703-711); per_image == 0: target is flat like pred. */
int fpg_l1_loss(const float* pred, const float* target, int64_t count, int64_t per_image,
                int64_t target_image_stride, float weight, float grad_scale, float* loss,
                float* dpred, int accumulate, float* scratch /* >= 512 floats */, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * fp32 parity mode (north_star: "fp32 rtol 1e-4"): every tensor-core operand is a pair of bf16 tensors hi + lo and a
 * convolution is hi_x*hi_w + lo_x*hi_w + hi_x*lo_w in ONE launch of the ordinary conv kernels -- activations hold the
 * channel blocks [hi | lo | hi] (3 * C channels), packed weights [hi_w | hi_w | lo_w] along the contraction dimension,
 * outputs stay fp32 (fpg_act.fp32 = FPG_DT_FP32). The two helpers below are the elementwise glue of that mode.
 * ---------------------------------------------------------------------------------------------------------- */
/* v = act(norm ? InstanceNorm(y) : y) [+ residual]; y: halo-free fp32 [n][h][w][c]; residual / skip_out (may be NULL):
 * dense fp32 [n][h][w][c]; out: bf16, 3 * c channels = [hi(v) | lo(v) | hi(v)], reflect halo out->halo mirrored.
 * (nn.InstanceNorm2d + relu / LeakyReLU / F.pad(reflect) / residual add of model_architectures.py:313-333,408-418.) */
int fpg_norm_split_f32(const fpg_act* y, int norm, float eps, int act, const float* residual, float* skip_out,
                       const fpg_act* out, void* stream);
/* src fp32 NCHW [n][c_src][h][w] -> dst bf16 [n][h+2halo][w+2halo][3 * cpad] = [hi | lo | hi], reflect halo, channels
 * >= c_src zero */
int fpg_pack_nchw_split(const float* src, int32_t c_src, const fpg_act* dst, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Layout / packing helpers around the network boundary (model.py:613-617: .to(device), torch.cat).
 * ---------------------------------------------------------------------------------------------------------- */
/* src fp32 NCHW [n][c_src][h][w] -> dst bf16 NHWC channels [c0, c0+c_src) of dst interior, reflect halo filled if
 * dst->halo > 0. Channels of dst outside the copied range are left untouched unless zero_rest != 0.
 * c_img (0 = c_src): channels per image of the tensor `src` points into, when the c_src channels are a slice of a wider
 * NCHW tensor (the topography conditions input_stack[:, 3:] re-attached to every synthetic image, model.py:683-689). */
int fpg_pack_nchw(const float* src, int32_t c_src, int32_t c_img, const fpg_act* dst, int32_t c0, int zero_rest,
                  void* stream);
/* The three packed copies of a paired training batch in one pass (train_paired, model.py:615-617): x fp32 NCHW
 * [n][c_x][h][w], y fp32 NCHW [n][c_y][h][w]  ->  gin (16-channel bf16, reflect halo: the generator input), fake and real
 * (16-channel bf16, no halo: the discriminator inputs torch.cat((x, .), 1); fake gets channels [0, c_x), its image
 * channels are written by fpg_blend_fwd; real gets x and y). Channels beyond the valid ones are 0. */
int fpg_pack_paired_inputs(const float* x, int32_t c_x, const float* y, int32_t c_y, const fpg_act* gin,
                           const fpg_act* fake, const fpg_act* real, void* stream);
/* Space-to-depth copy for a 4x4 stride-2 pad-1 convolution over a 16-channel tensor (PatchGAN model.0,
 * model_architectures.py:424): dst [n][h/2+1][w/2+1][64] bf16, block (by, bx) = pixels (2by-1+i, 2bx-1+j) of src as
 * channels (2i+j)*16 + c, zeros outside the image. The convolution then is 2x2 stride-1 pad-0 over 64 channels with
 * weights W2[k][ty][tx][(2i+j)*16 + c] = W[k][c][2ty+i][2tx+j] (a tap permutation of the ordinary fprop operand). */
int fpg_space_to_depth16(const fpg_act* src, const fpg_act* dst, void* stream);
/* Backward through a fused tanh head (CycleGAN generator, model_architectures.py:115-116): dpre (bf16 NHWC) =
 * dout (fp32 NCHW [n][c_valid][h][w]) * (1 - out^2), out = the fp32 NHWC tanh output; channels >= c_valid are 0. */
int fpg_tanh_bwd_pack(const float* dout_nchw, const fpg_act* out, int32_t c_valid, const fpg_act* dpre, void* stream);
/* src (bf16 or fp32 per src->fp32) NHWC channels [c0, c0+c_dst) of the interior -> dst fp32 NCHW; accumulate != 0
 * adds instead */
int fpg_unpack_nchw(const fpg_act* src, int32_t c0, float* dst, int32_t c_dst, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Optimiser -- torch.optim.Adam(lr, betas=(0.5, 0.999), eps=1e-8), model.py:119-122, one launch per flat buffer.
 *   p, g, m, v: fp32 [count]; step is the 1-based step count; grad_scale multiplies g first (1/world_size).
 * ---------------------------------------------------------------------------------------------------------- */
int fpg_adam_step(float* p, const float* g, float* m, float* v, int64_t count, float lr, float beta1, float beta2,
                  float eps, int32_t step, float grad_scale, void* stream);

/* Same update with the step count, learning rate and bias corrections in DEVICE memory, so that the launch
 * arguments never change and the whole training step can be replayed as a CUDA graph.
 *   state: int32[4] = {step (incremented by this call), lr as float bits, scratch, scratch}. */
int fpg_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t count, float beta1, float beta2, float eps,
                      int32_t* state, float grad_scale, void* stream);

/* Adam AND the repack of the bf16 GEMM operands in one launch (csrc/adam_pack.cu): replaces fpg_adam_step_dev followed
 * by fpg_pack_weights_batched after every optimiser step (model.py:633,646). The flat buffers p / m / v / gradient
 * source(s) hold every parameter of one optimiser; tables in device memory describe them:
 *   layers[i]  a convolution weight [k][c][rs] at element offset p_off of the flat buffers, cut into tiles of tk x tc
 *              (k, c) pairs (fpg_adam_pack_tile gives the tile for a tap count); job[0..n_jobs <= 8) index `jobs`: the
 *              fpg_pack_job descriptors of its operands as built by fpg_pack_jobs (their `src` is not used). Padding
 *              rows / taps / columns of the operands are NOT rewritten: pack once with fpg_pack_weights_batched first;
 *   chunks[i]  count <= 4096 other parameters at element offset off; copy_dst (optional) receives the new values
 *              (the zero-padded fp32 bias vector of a layer whose epilogue adds a bias);
 *   block b    works on tile block_first[b] of layer block_item[b] (>= 0), or on chunk -1 - block_item[b].
 * grads_host: n_src <= 16 device pointers, summed in that order (data-parallel peer exchange); gsum_out (optional,
 * may alias a source) receives the sum. state as in fpg_adam_step_dev. */
typedef struct {
  int64_t p_off;
  int32_t k, c, rs;
  int32_t tk, tc;
  int32_t n_jobs;
  int32_t job[8];
} fpg_adam_pack_layer;
typedef struct {
  int64_t off;
  float* copy_dst;
  int32_t count;
  int32_t pad_;
} fpg_adam_pack_chunk;
int fpg_adam_pack_tile(int32_t rs, int32_t* tk, int32_t* tc);
int fpg_adam_pack_step(float* p, const void* const* grads_host, int32_t n_src, float* m, float* v, float beta1,
                       float beta2, float eps, int32_t* state, float grad_scale, float* gsum_out,
                       const fpg_adam_pack_layer* layers_dev, const fpg_pack_job* jobs_dev,
                       const fpg_adam_pack_chunk* chunks_dev, const int32_t* block_item_dev,
                       const int32_t* block_first_dev, int32_t n_blocks, void* stream);

/* The first half of fpg_adam_step_dev alone: advance state[0] and refresh the bias-correction scalars. */
int fpg_adam_prepare_dev(int32_t* state, float beta1, float beta2, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Peer exchange -- data-parallel gradient sum over NVLink / NVSwitch peer memory, no reference counterpart (the
 * reference is one process; this is what stands between loss.backward() and optimizer.step(), model.py:632-633 and
 * :645-646, when the batch is sharded over W GPUs). Copy engines push each rank's gradient buckets into a staging slot
 * on every peer (no SM involved, overlaps the backward pass); Adam sums the W sources in rank order (bit-identical
 * replicas, equal to one process accumulating the shards in order). See csrc/peer.cu.
 *   fpg_peer_alloc / free      a zero-filled device block of its own cudaMalloc (IPC handles address whole blocks);
 *   fpg_peer_export            handle_host <- the FPG_PEER_HANDLE_BYTES-byte IPC handle of such a block;
 *   fpg_peer_open / close      map / unmap another process's block (peer access enabled on first use);
 *   fpg_peer_copy              asynchronous device-to-device copy on `stream` (copy engine; capturable);
 *   fpg_peer_signal            *flags_host[i] = *ctr + add for i < n (system-scope release stores into peer memory,
 *                              ordered after everything earlier on the stream), then *ctr += bump;
 *   fpg_peer_wait              block the stream until local_flags[i] >= *ctr + add for every i < n except i == skip
 *                              (words the peers write); status = device int64[2]: after timeout_s seconds status[0] =
 *                              i + 1 and the kernel traps; status[1] accumulates the nanoseconds spent waiting;
 *   fpg_peer_push              src[0, bytes) -> every dst_host[i] (i < n_dst <= 16, peer memory) by ONE kernel (16-byte
 *                              accesses): for the tail of an exchange, where nothing is left to overlap with;
 *   fpg_adam_step_dev_multi    fpg_adam_step_dev on g = ((grads_host[0] + grads_host[1]) + ...) (n_src <= 16 fp32
 *                              device buffers, summed in that order); gsum_out (optional, may alias a source) <- g.
 * ---------------------------------------------------------------------------------------------------------- */
#define FPG_PEER_HANDLE_BYTES 64
int fpg_peer_alloc(void** ptr, int64_t bytes);
int fpg_peer_free(void* ptr);
int fpg_peer_export(void* ptr, void* handle_host);
int fpg_peer_open(const void* handle_host, void** ptr);
int fpg_peer_close(void* ptr);
int fpg_peer_copy(void* dst, const void* src, int64_t bytes, void* stream);
int fpg_peer_signal(void* const* flags_host, int32_t n, uint32_t* ctr, uint32_t add, uint32_t bump, void* stream);
int fpg_peer_wait(const uint32_t* local_flags, int32_t n, int32_t skip, const uint32_t* ctr, uint32_t add,
                  int64_t* status, float timeout_s, void* stream);
int fpg_peer_push(const void* src, void* const* dst_host, int32_t n_dst, int64_t bytes, void* stream);
int fpg_adam_step_dev_multi(float* p, const void* const* grads_host, int32_t n_src, float* m, float* v, int64_t count,
                            float beta1, float beta2, float eps, int32_t* state, float grad_scale, float* gsum_out,
                            void* stream);

/* dst += src (fp32, 16-byte aligned): accumulates the parameter gradients of a network that runs several times in one
 * training step (train_cycle applies each generator 2-3 times, model.py:683-704). */
int fpg_add_f32(float* dst, const float* src, int64_t count, void* stream);

/* Device-resident history buffer of generated images (get_buffer_image, model.py:275-294): pool holds 50 entries of
 * entry_bytes; ctrl = device int32[2] {use_slot, store_slot}, -1 = none, decided by the host's random draws.
 * out = use_slot >= 0 ? pool[use_slot] : cur; then pool[store_slot] = cur if store_slot >= 0 (same slot = exchange). */
int fpg_history_exchange(const void* cur, void* pool, const int32_t* ctrl, void* out, int64_t entry_bytes,
                         void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Flood-mask thresholding -- (sigmoid(logit) > 0.5).float(), model.py:399-400, segmentation_model.py:244-248.
 * Bit-exact with the fp32 reference expression (sigmoid evaluated in fp32 as 1/(1+exp(-x)), then compared).
 * ---------------------------------------------------------------------------------------------------------- */
int fpg_flood_mask(const float* logits, float* mask, int64_t count, void* stream);
/* confusion counts {tp, fp, tn, fn} of pred vs truth masks (values 0/1), as int64[4] */
int fpg_confusion_counts(const float* pred, const float* truth, int64_t count, int64_t* counts4, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FPG_H_ */
