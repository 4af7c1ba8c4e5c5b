// Planners: convolution geometry -> implicit-GEMM descriptors, plus weight packing and the split-K wgrad reduce.
// Pure host logic except for the small pack / reduce kernels at the bottom.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host_util.h"

namespace fpg {

static int pick_cblk(int c) {
  if (c % 64 == 0) return 64;
  if (c == 32) return 32;
  if (c == 16) return 16;
  return -1;
}

static int pow2_at_least(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// 5-D TMA view (c, x, plane, y, n) of an activation buffer (interior + halo). stride 2 folds the x parity into
// dim 0 and the y parity into dim 2 so that a box of consecutive coordinates walks every second pixel.
static void make_act_view(const fpg_act* a, int stride, int box_c, int tile_w, int tile_h, fpg_tmap* m) {
  const uint64_t hp = a->h + 2 * a->halo, wp = a->w + 2 * a->halo, cs = a->c_stride;
  memset(m, 0, sizeof(*m));
  m->base = a->data;
  m->rank = 5;
  m->swizzle_bytes = box_c * 2;
  if (stride == 1) {
    m->dims[0] = a->c;
    m->dims[1] = wp;
    m->dims[2] = 1;
    m->dims[3] = hp;
    m->dims[4] = a->n;
    m->strides[0] = cs * 2;
    m->strides[1] = wp * cs * 2;
    m->strides[2] = wp * cs * 2;
    m->strides[3] = hp * wp * cs * 2;
  } else {
    m->dims[0] = cs + a->c;
    m->dims[1] = wp / 2;
    m->dims[2] = 2;
    m->dims[3] = hp / 2;
    m->dims[4] = a->n;
    m->strides[0] = 2 * cs * 2;
    m->strides[1] = wp * cs * 2;
    m->strides[2] = 2 * wp * cs * 2;
    m->strides[3] = hp * wp * cs * 2;
  }
  m->box[0] = box_c;
  m->box[1] = tile_w;
  m->box[2] = 1;
  m->box[3] = tile_h;
  m->box[4] = 1;
}

// tap (r, s) of a forward conv reading input pixel (o*stride + r - pad) as view coordinates
static fpg_tap fwd_tap(int r, int s, int stride, int pad, int cs) {
  fpg_tap t;
  if (stride == 1) {
    t.c0 = 0;
    t.dx = s - pad;
    t.plane = 0;
    t.dy = r - pad;
  } else {
    const int qx = s - pad, qy = r - pad;
    t.c0 = pos_mod(qx, 2) * cs;
    t.dx = floor_div(qx, 2);
    t.plane = pos_mod(qy, 2);
    t.dy = floor_div(qy, 2);
  }
  return t;
}

static void out_view_of(const fpg_act* y, int with_halo, fpg_out_view* o) {
  const int64_t wp = y->w + 2 * y->halo, hp = y->h + 2 * y->halo, cs = y->c_stride;
  const int64_t org = with_halo ? 0 : (static_cast<int64_t>(y->halo) * wp + y->halo) * cs;
  o->base = y->fp32 == FPG_DT_FP32 ? static_cast<void*>(static_cast<float*>(y->data) + org)
                    : static_cast<void*>(static_cast<__nv_bfloat16*>(y->data) + org);
  o->stride_n = hp * wp * cs;
  o->stride_y = wp * cs;
  o->stride_x = cs;
  o->mul_y = o->mul_x = 1;
  o->off_y = o->off_x = 0;
  o->valid_h = with_halo ? static_cast<int>(hp) : y->h;
  o->valid_w = with_halo ? static_cast<int>(wp) : y->w;
  o->fp32 = y->fp32;
}

// Pixel tile (tile_w x tile_h == pixels, tile_w a power of two) covering a ho x wo grid with the fewest tiles;
// ties go to the widest tile (longest contiguous runs per TMA box row).
static void pick_tile(int ho, int wo, int* tile_w, int* tile_h, int pixels) {
  int best_tw = pixels, best_tiles = -1;
  for (int tw = pixels; tw >= 1; tw >>= 1) {
    const int th = pixels / tw;
    if (th > 256) break;  // TMA box dimension limit
    const int tiles = ceil_div(wo, tw) * ceil_div(ho, th);
    if (best_tiles < 0 || tiles < best_tiles) {
      best_tiles = tiles;
      best_tw = tw;
    }
  }
  *tile_w = best_tw;
  *tile_h = pixels / best_tw;
}

static int pick_block_n(int n_total, int m_tiles, int sms) {
  int bn = n_total <= 256 ? n_total : 256;
  while (n_total % bn != 0) bn -= 16;
  while (m_tiles * (n_total / bn) < sms && bn % 32 == 0 && bn > 32) bn /= 2;
  return bn;
}

// Chooses the 128-pixel CTA tile shape and block_n of the fprop-type kernel. (A 256-pixel single-CTA tile -- two MMAs
// sharing one B tile -- was measured slower on B200: its accumulators fill TMEM, so the epilogue is exposed; the 2-CTA
// kernel is how operand traffic is cut instead, see maybe_pair.)
static void choose_fprop_tile(int ho, int wo, int n_img, int n_total, int cblk, int sms, int* tile_w, int* tile_h,
                              int* block_n) {
  int tw1, th1;
  pick_tile(ho, wo, &tw1, &th1, 128);
  const int t1 = n_img * ceil_div(wo, tw1) * ceil_div(ho, th1);
  const int bn1 = pick_block_n(n_total, t1, sms);
  *tile_w = tw1;
  *tile_h = th1;
  *block_n = bn1;
}

// 2-CTA (cta_group::2) variant of a planned fprop-type launch: pair tiles of 2*tile_h x tile_w pixels, each CTA loads
// half of the weight tile. Chosen when there is enough work for the 74 clusters and the layer is wide enough to be
// bound by operand delivery.
static void maybe_pair(fpg_igemm_fprop_desc* d, int grid_h, int sms) {
  d->cta_pair = 0;
  if (getenv("FPG_DISABLE_2CTA") != nullptr) return;
  if (d->cblk != 64 || d->block_n < 128 || d->block_n % 32 != 0 || d->tile_h * d->tile_w != 128) return;
  const int single = d->n_img * d->tiles_x * d->tiles_y * d->n_blocks;
  const int pair_rows = ceil_div(grid_h, 2 * d->tile_h);
  const int pairs = d->n_img * d->tiles_x * pair_rows * d->n_blocks;
  const double cost_single = static_cast<double>(ceil_div(single, sms));
  const double cost_pair = static_cast<double>(ceil_div(pairs, sms / 2)) * 0.88;  // measured pair/single tile time
  if (cost_pair >= cost_single) return;
  d->cta_pair = 1;
  d->tiles_y = pair_rows;
  d->b.box[1] = d->block_n / 2;
  int s2 = (200 * 1024) / (16384 + d->block_n * 64);
  d->stages = s2 > 8 ? 8 : s2;
}

static int pick_stages(int stage_bytes) {
  int s = (200 * 1024) / stage_bytes;
  if (s > 8) s = 8;
  if (s < 2) s = 2;
  return s;
}

// Taps of dgrad parity class (pi, pj) for a forward conv g, in the order used by both the plan and the packing.
// Returns the number of taps; rs[i] = r*S+s of the forward filter tap, taps[i] = coordinates into dy (stride-1 view).
static int dgrad_class_taps(const fpg_conv_geom* g, int pi, int pj, int* rs, fpg_tap* taps) {
  int n = 0;
  for (int r = 0; r < g->r; ++r) {
    if (g->stride == 2 && pos_mod(pi + g->pad - r, 2) != 0) continue;
    for (int s = 0; s < g->s; ++s) {
      if (g->stride == 2 && pos_mod(pj + g->pad - s, 2) != 0) continue;
      fpg_tap t;
      t.c0 = 0;
      t.plane = 0;
      if (g->stride == 1) {
        t.dy = g->pad - r;
        t.dx = g->pad - s;
      } else {
        t.dy = (pi + g->pad - r) / 2;  // exact
        t.dx = (pj + g->pad - s) / 2;
      }
      rs[n] = r * g->s + s;
      taps[n] = t;
      ++n;
    }
  }
  return n;
}

static int padded_taps(int ntaps, int c, int cblk) {
  const int sub_per_stage = 64 / cblk;
  const int chunks = c / cblk;
  int t = ntaps;
  while ((t * chunks) % sub_per_stage != 0) ++t;
  return t;
}

static int plan_fprop(const fpg_act* x, const void* w, const float* bias, int act, const fpg_conv_geom* g,
                      const fpg_act* y, int sms, fpg_igemm_fprop_desc* d) {
  FPG_REQUIRE(x && g && y && d, "null argument");
  FPG_REQUIRE(x->fp32 == FPG_DT_BF16, "conv input must be bf16 (the tensor-core operand type)");
  FPG_REQUIRE(g->stride == 1 || g->stride == 2, "stride %d", g->stride);
  FPG_REQUIRE(x->halo == 0 || g->pad == 0, "input halo %d with zero pad %d", x->halo, g->pad);
  FPG_REQUIRE(x->c == g->c_in && y->c == g->c_out, "channels x %d/%d y %d/%d", x->c, g->c_in, y->c, g->c_out);
  const int hp = x->h + 2 * x->halo, wp = x->w + 2 * x->halo;
  const int ho = (hp + 2 * g->pad - g->r) / g->stride + 1, wo = (wp + 2 * g->pad - g->s) / g->stride + 1;
  FPG_REQUIRE(ho == y->h && wo == y->w && x->n == y->n, "output %dx%d expected %dx%d", y->h, y->w, ho, wo);
  FPG_REQUIRE(g->stride == 1 || (hp % 2 == 0 && wp % 2 == 0), "stride-2 input must have even extent");
  const int cblk = pick_cblk(g->c_in);
  FPG_REQUIRE(cblk > 0 && g->c_out % 16 == 0, "unsupported channels c_in %d c_out %d", g->c_in, g->c_out);
  memset(d, 0, sizeof(*d));
  d->cblk = cblk;
  d->c_per_tap = g->c_in;
  const int real_taps = g->r * g->s;
  d->num_taps = padded_taps(real_taps, g->c_in, cblk);
  FPG_REQUIRE(d->num_taps <= FPG_MAX_TAPS, "too many taps %d", d->num_taps);
  d->num_sub = d->num_taps * (g->c_in / cblk);
  for (int r = 0; r < g->r; ++r)
    for (int s = 0; s < g->s; ++s) d->taps[r * g->s + s] = fwd_tap(r, s, g->stride, g->pad, x->c_stride);
  for (int t = real_taps; t < d->num_taps; ++t) d->taps[t] = d->taps[0];
  choose_fprop_tile(ho, wo, x->n, g->c_out, cblk, sms, &d->tile_w, &d->tile_h, &d->block_n);
  d->tiles_x = ceil_div(wo, d->tile_w);
  d->tiles_y = ceil_div(ho, d->tile_h);
  d->n_img = x->n;
  d->n_blocks = g->c_out / d->block_n;
  d->stages = pick_stages(d->tile_w * d->tile_h * 128 + d->block_n * 128);
  d->act = act;
  d->bias = bias;
  make_act_view(x, g->stride, cblk, d->tile_w, d->tile_h, &d->a);
  memset(&d->b, 0, sizeof(d->b));
  d->b.base = const_cast<void*>(w);
  d->b.rank = 2;
  d->b.swizzle_bytes = cblk * 2;
  d->b.dims[0] = static_cast<uint64_t>(d->num_sub) * cblk;
  d->b.dims[1] = g->c_out;
  d->b.strides[0] = d->b.dims[0] * 2;
  d->b.box[0] = cblk;
  d->b.box[1] = d->block_n;
  out_view_of(y, 0, &d->out);
  maybe_pair(d, ho, sms);
  return 0;
}

// K extent (elements) of dgrad class q's packed matrix, and its element offset inside the packed buffer
static void dgrad_class_layout(const fpg_conv_geom* g, int64_t* k_of, int64_t* off_of, int* nclasses) {
  const int cblk = pick_cblk(g->c_out);
  const int nc = g->stride == 2 ? 4 : 1;
  int64_t off = 0;
  for (int q = 0; q < nc; ++q) {
    int rs[FPG_MAX_TAPS];
    fpg_tap taps[FPG_MAX_TAPS];
    const int nt = dgrad_class_taps(g, q >> 1, q & 1, rs, taps);
    const int ntp = padded_taps(nt, g->c_out, cblk);
    k_of[q] = static_cast<int64_t>(ntp) * g->c_out;
    off_of[q] = off;
    off += k_of[q] * g->c_in;
  }
  *nclasses = nc;
}

static int plan_dgrad(const fpg_act* dy, const void* wt, const float* bias, int act, const fpg_conv_geom* g,
                      const fpg_act* dx, int sms, fpg_igemm_fprop_desc* descs, int* n_descs) {
  FPG_REQUIRE(dy && g && dx && descs && n_descs, "null argument");
  FPG_REQUIRE(dy->fp32 == FPG_DT_BF16, "dgrad input must be bf16 (the tensor-core operand type)");
  FPG_REQUIRE(g->stride == 1 || g->stride == 2, "stride %d", g->stride);
  FPG_REQUIRE(dy->halo == 0, "dy must not have a halo");
  FPG_REQUIRE(dx->halo == 0 || g->pad == 0, "dx halo %d with zero pad %d", dx->halo, g->pad);
  FPG_REQUIRE(dy->c == g->c_out && dx->c == g->c_in, "channels dy %d/%d dx %d/%d", dy->c, g->c_out, dx->c, g->c_in);
  const int hp = dx->h + 2 * dx->halo, wp = dx->w + 2 * dx->halo;
  const int ho = (hp + 2 * g->pad - g->r) / g->stride + 1, wo = (wp + 2 * g->pad - g->s) / g->stride + 1;
  FPG_REQUIRE(ho == dy->h && wo == dy->w && dx->n == dy->n, "dy %dx%d expected %dx%d", dy->h, dy->w, ho, wo);
  FPG_REQUIRE(g->stride == 1 || (hp % 2 == 0 && wp % 2 == 0), "stride-2 dx must have even extent");
  const int cblk = pick_cblk(g->c_out);
  FPG_REQUIRE(cblk > 0 && g->c_in % 16 == 0, "unsupported channels c_in %d c_out %d", g->c_in, g->c_out);
  int64_t k_of[4], off_of[4];
  int nc = 0;
  dgrad_class_layout(g, k_of, off_of, &nc);
  for (int q = 0; q < nc; ++q) {
    fpg_igemm_fprop_desc* d = &descs[q];
    memset(d, 0, sizeof(*d));
    const int pi = q >> 1, pj = q & 1;
    int rs[FPG_MAX_TAPS];
    const int nt = dgrad_class_taps(g, pi, pj, rs, d->taps);
    FPG_REQUIRE(nt > 0, "empty dgrad parity class");
    d->cblk = cblk;
    d->c_per_tap = g->c_out;
    d->num_taps = padded_taps(nt, g->c_out, cblk);
    FPG_REQUIRE(d->num_taps <= FPG_MAX_TAPS, "too many taps %d", d->num_taps);
    for (int t = nt; t < d->num_taps; ++t) d->taps[t] = d->taps[0];
    d->num_sub = d->num_taps * (g->c_out / cblk);
    const int oh = g->stride == 2 ? hp / 2 : hp, ow = g->stride == 2 ? wp / 2 : wp;  // class output grid
    choose_fprop_tile(oh, ow, dx->n, g->c_in, cblk, sms, &d->tile_w, &d->tile_h, &d->block_n);
    d->tiles_x = ceil_div(ow, d->tile_w);
    d->tiles_y = ceil_div(oh, d->tile_h);
    d->n_img = dx->n;
    d->n_blocks = g->c_in / d->block_n;
    d->stages = pick_stages(d->tile_w * d->tile_h * 128 + d->block_n * 128);
    d->act = act;
    d->bias = bias;
    make_act_view(dy, 1, cblk, d->tile_w, d->tile_h, &d->a);
    d->b.base = static_cast<void*>(static_cast<__nv_bfloat16*>(const_cast<void*>(wt)) + off_of[q]);
    d->b.rank = 2;
    d->b.swizzle_bytes = cblk * 2;
    d->b.dims[0] = static_cast<uint64_t>(k_of[q]);
    d->b.dims[1] = g->c_in;
    d->b.strides[0] = d->b.dims[0] * 2;
    d->b.box[0] = cblk;
    d->b.box[1] = d->block_n;
    out_view_of(dx, 1, &d->out);
    if (g->stride == 2) {
      d->out.mul_y = d->out.mul_x = 2;
      d->out.off_y = pi;
      d->out.off_x = pj;
      d->out.valid_h = oh;
      d->out.valid_w = ow;
    }
    // An extent just above a multiple of 64 (66 = 64 + reflect halo) quantises badly into one tile shape (16 x 8
    // tiles cover 80 x 72 of it): cover x < 64k with 64 x 2 tiles and the few edge columns with tall narrow tiles.
    const int rem = ow % 64;
    if (g->stride == 1 && ow > 64 && rem > 0 && rem <= 8 && getenv("FPG_DISABLE_TILE_REGIONS") == nullptr) {
      const int tw1 = pow2_at_least(rem), th1 = 128 / tw1;
      const int t0 = (ow / 64) * ceil_div(oh, 2), t1 = ceil_div(oh, th1);
      const int single = d->tiles_x * d->tiles_y;
      const int nb_single = d->n_blocks;
      const int bn2 = pick_block_n(g->c_in, dx->n * (t0 + t1), sms);
      const int waves_single = ceil_div(dx->n * single * nb_single, sms);
      const int waves_two = ceil_div(dx->n * (t0 + t1) * (g->c_in / bn2), sms);
      if (waves_two * bn2 < waves_single * d->block_n) {  // compare in units of (waves x N columns per tile)
        d->tile_w = 64;
        d->tile_h = 2;
        d->tiles_x = ow / 64;
        d->tiles_y = ceil_div(oh, 2);
        d->block_n = bn2;
        d->n_blocks = g->c_in / bn2;
        d->b.box[1] = bn2;
        d->stages = pick_stages(128 * 128 + bn2 * 128);
        make_act_view(dy, 1, cblk, 64, 2, &d->a);
        make_act_view(dy, 1, cblk, tw1, th1, &d->a1);
        d->tile_w1 = tw1;
        d->tile_h1 = th1;
        d->tiles_x1 = 1;
        d->tiles_y1 = t1;
        d->x_org1 = (ow / 64) * 64;
        continue;  // regions are not combined with CTA pairs
      }
    }
    maybe_pair(d, oh, sms);
  }
  *n_descs = nc;
  return 0;
}

// Row-stationary plan (igemm_rows.cu). Returns 0 when planned, 1 when the path does not apply.
static int plan_rows(const fpg_act* a, const void* w, const float* bias, int act, const fpg_conv_geom* g,
                     const fpg_act* out, int dgrad, int sms, fpg_igemm_rows_desc* d) {
  FPG_REQUIRE(a && g && out && d, "null argument");
  FPG_REQUIRE(a->fp32 == FPG_DT_BF16, "conv input must be bf16 (the tensor-core operand type)");
  if (getenv("FPG_DISABLE_ROWS") != nullptr) return 1;
  if (g->stride != 1 || g->r * g->s <= 1 || g->r * g->s > FPG_MAX_TAPS) return 1;
  const int ca = dgrad ? g->c_out : g->c_in;     // channels of the gathered operand
  const int n_total = dgrad ? g->c_in : g->c_out;  // channels produced
  if (!(ca == 16 || ca == 32 || ca == 64) || n_total > 128 || n_total % 16 != 0) return 1;
  FPG_REQUIRE(a->c == ca && out->c == n_total, "channels a %d/%d out %d/%d", a->c, ca, out->c, n_total);
  int oh, ow;  // output grid
  if (!dgrad) {
    FPG_REQUIRE(a->halo == 0 || g->pad == 0, "input halo %d with zero pad %d", a->halo, g->pad);
    const int hp = a->h + 2 * a->halo, wp = a->w + 2 * a->halo;
    oh = hp + 2 * g->pad - g->r + 1;
    ow = wp + 2 * g->pad - g->s + 1;
    FPG_REQUIRE(oh == out->h && ow == out->w && a->n == out->n, "output %dx%d expected %dx%d", out->h, out->w, oh, ow);
  } else {
    FPG_REQUIRE(a->halo == 0, "dy must not have a halo");
    FPG_REQUIRE(out->halo == 0 || g->pad == 0, "dx halo %d with zero pad %d", out->halo, g->pad);
    oh = out->h + 2 * out->halo;
    ow = out->w + 2 * out->halo;
    FPG_REQUIRE(oh + 2 * g->pad - g->r + 1 == a->h && ow + 2 * g->pad - g->s + 1 == a->w && a->n == out->n,
                "dy %dx%d does not match dx", a->h, a->w);
  }
  if (ow < 96) return 1;
  memset(d, 0, sizeof(*d));
  d->cblk = ca;
  d->block_n = n_total;
  d->rows = g->r;
  d->cols = g->s;
  int th = 256 / n_total;
  if (th > 4) th = 4;
  const int64_t tap_bytes = static_cast<int64_t>(n_total) * ca * 2;
  const int64_t b_row = tap_bytes * g->s;
  const int64_t a_slot = ((static_cast<int64_t>(128 + g->s - 1) * ca * 2) + 1023) & ~static_cast<int64_t>(1023);
  const int64_t budget = 222 * 1024;  // of the 227 KB a CTA may use (barriers + 1 KB alignment slack come on top)
  if (b_row * g->r + 3 * a_slot <= budget) {
    d->b_stages = g->r;  // the whole filter stays in shared memory
  } else {
    // a filter row is released th - 1 steps after its first use: th + 1 slots prefetch ONE row ahead, which does not
    // cover the L2 latency of a 28 KB row inside one ~0.7 us step; th + 2 does
    while (th > 1 && b_row * (th + 1) + 3 * a_slot > budget) --th;
    if (b_row * (th + 1) + 3 * a_slot > budget) return 1;
    d->b_stages = b_row * (th + 2) + 3 * a_slot <= budget ? th + 2 : th + 1;
  }
  d->tile_rows = th;
  int64_t as = (budget - b_row * d->b_stages) / a_slot;
  d->a_stages = as > 8 ? 8 : static_cast<int>(as);
  const int real_taps = g->r * g->s;
  if (!dgrad) {
    d->dy0 = -g->pad;
    d->dx0 = -g->pad;
    for (int i = 0; i < g->r; ++i)
      for (int j = 0; j < g->s; ++j) d->tap_of[i * g->s + j] = static_cast<int16_t>(i * g->s + j);
    d->b.dims[0] = static_cast<uint64_t>(padded_taps(real_taps, ca, ca)) * ca;
    d->b.base = const_cast<void*>(w);
    out_view_of(out, 0, &d->out);
  } else {
    // dx[p] = sum_{r,s} dy[p + pad - (r,s)] * w[r,s]: ascending offsets run through the filter backwards
    d->dy0 = g->pad - (g->r - 1);
    d->dx0 = g->pad - (g->s - 1);
    for (int i = 0; i < g->r; ++i)
      for (int j = 0; j < g->s; ++j)
        d->tap_of[i * g->s + j] = static_cast<int16_t>((g->r - 1 - i) * g->s + (g->s - 1 - j));
    int64_t k_of[4], off_of[4];
    int nc = 0;
    dgrad_class_layout(g, k_of, off_of, &nc);
    d->b.dims[0] = static_cast<uint64_t>(k_of[0]);
    d->b.base = const_cast<void*>(w);
    out_view_of(out, 1, &d->out);
  }
  d->b.rank = 2;
  d->b.swizzle_bytes = ca * 2;
  d->b.dims[1] = n_total;
  d->b.strides[0] = d->b.dims[0] * 2;
  d->b.box[0] = ca;
  d->b.box[1] = n_total;
  make_act_view(a, 1, ca, 128 + g->s - 1, 1, &d->a);
  d->n_img = a->n;
  d->tiles_x = ceil_div(ow, 128);
  d->tiles_y = ceil_div(oh, th);
  d->act = act;
  d->bias = bias;
  (void)sms;
  return 0;
}

static int largest_divisor_le(int v, int cap) {
  for (int d = cap; d >= 1; --d)
    if (v % d == 0) return d;
  return 1;
}

// width efficiency of 64 x 1 k tiles on a row of `w` pixels: shifted operands need one-row tiles
static bool rows_of_64_ok(int w) { return 10 * w >= 7 * 64 * ceil_div(w, 64); }  // >= 70 % of the tile pixels used

static int plan_wgrad(const fpg_act* x, const fpg_act* dy, const fpg_conv_geom* g, int sms, fpg_igemm_wgrad_desc* d) {
  FPG_REQUIRE(x && dy && g && d, "null argument");
  FPG_REQUIRE(x->fp32 == FPG_DT_BF16 && dy->fp32 == FPG_DT_BF16, "wgrad operands must be bf16");
  FPG_REQUIRE(g->stride == 1 || g->stride == 2, "stride %d", g->stride);
  FPG_REQUIRE(dy->halo == 0, "dy must not have a halo");
  FPG_REQUIRE(x->halo == 0 || g->pad == 0, "input halo %d with zero pad %d", x->halo, g->pad);
  FPG_REQUIRE(x->c == g->c_in && dy->c == g->c_out, "channels x %d/%d dy %d/%d", x->c, g->c_in, dy->c, g->c_out);
  const int hp = x->h + 2 * x->halo, wp = x->w + 2 * x->halo;
  const int ho = (hp + 2 * g->pad - g->r) / g->stride + 1, wo = (wp + 2 * g->pad - g->s) / g->stride + 1;
  FPG_REQUIRE(ho == dy->h && wo == dy->w && x->n == dy->n, "dy %dx%d expected %dx%d", dy->h, dy->w, ho, wo);
  FPG_REQUIRE(g->stride == 1 || (hp % 2 == 0 && wp % 2 == 0), "stride-2 input must have even extent");
  const int ntaps = g->r * g->s;
  FPG_REQUIRE(ntaps <= FPG_MAX_TAPS, "too many taps");
  memset(d, 0, sizeof(*d));
  d->taps_r = g->r;
  d->taps_s = g->s;
  d->y_shifts = 1;
  pick_tile(ho, wo, &d->tile_w, &d->tile_h, 64);
  if (getenv("FPG_EXP_WGRAD_ROW_TILES") != nullptr) {  // experiment: one-row k tiles without shifted operands
    d->tile_w = 64;
    d->tile_h = 1;
  }
  d->kt_x = ceil_div(wo, d->tile_w);
  d->kt_y = ceil_div(ho, d->tile_h);
  d->n_img = x->n;
  fpg_tap null_tap = {0, 0, 0, 0};
  fpg_tap in_taps[FPG_MAX_TAPS];
  for (int r = 0; r < g->r; ++r)
    for (int s = 0; s < g->s; ++s) in_taps[r * g->s + s] = fwd_tap(r, s, g->stride, g->pad, x->c_stride);
  const bool shift_ok = getenv("FPG_DISABLE_WGRAD_SHIFT") == nullptr && g->stride == 1 && g->s >= 2;
  int x_extra = 0, y_extra = 0;  // additional pixels of the X / Y boxes (shifted operands)

  const bool big_out = g->c_out % 64 == 0, big_in = g->c_in % 64 == 0;
  // measured on B200: not faster than the M = 256 / N = 256 plan below (N = 128 MMAs read 8 KB of operands per 64
  // cycles, which together with the TMA writes saturates the 128 B/clk of shared memory): opt-in for experiments
  const bool wide_shift = shift_ok && getenv("FPG_WGRAD_SHIFT_WIDE") != nullptr;
  if (big_out && big_in && wide_shift && g->c_out % 128 == 0 && g->c_in % 128 == 0 && g->s * 128 <= 512 &&
      wo % 64 == 0) {
    // Wide layers on 64-pixel rows (residual 3x3 convs): M = 128 output channels, N = 128 input channels, the S column
    // taps of a filter row are S MMA groups reading ONE (64 + S - 1)-pixel input box at shifted pixel rows.
    d->x_is_dy = 1;
    d->x_ca = 64;
    d->x_atoms = 2;
    d->x_groups = g->c_out / 128;
    d->x_taps_mode = 0;
    d->x_ntaps = 1;
    d->x_taps[0] = null_tap;
    d->y_ca = 64;
    d->y_atoms = 2;
    d->y_groups = g->c_in / 128;
    d->y_taps_mode = 0;
    d->y_ntaps = g->r;
    d->y_shifts = g->s;
    for (int r = 0; r < g->r; ++r) {
      d->y_taps[r] = in_taps[r * g->s];
      d->y_tap_rs[r] = static_cast<int16_t>(r * g->s);
    }
    d->tile_w = 64;
    d->tile_h = 1;
    d->kt_x = wo / 64;
    d->kt_y = ho;
    y_extra = g->s - 1;
    make_act_view(dy, 1, 64, 64, 1, &d->x);
    make_act_view(x, 1, 64, 64 + y_extra, 1, &d->y);
  } else if (big_out && !big_in && shift_ok && g->c_out == 64 && g->s * g->c_in <= 256 && rows_of_64_ok(wo)) {
    // Few input channels (generator stem): M = 128 = dy at output rows o - 1 and o (two filter rows per item),
    // N = S x c_in = the S column taps as pixel-shift atoms of one input box. Output rows run from -1 so that the
    // second atom sees every row of dy.
    FPG_REQUIRE(g->c_in == 16 || g->c_in == 32, "unsupported c_in %d", g->c_in);
    d->x_is_dy = 1;
    d->x_ca = 64;
    d->x_atoms = 2;
    d->x_groups = 1;
    d->x_taps_mode = 1;
    d->x_ntaps = 2;
    fpg_tap up = {0, 0, 0, -1};
    d->x_taps[0] = up;        // dy[o - 1 + ...]: pairs with input row (o - 1) + r_y  -> filter row r_y
    d->x_taps[1] = null_tap;  // dy[o]:            pairs with input row (o - 1) + r_y  -> filter row r_y - 1
    d->x_tap_rs[0] = 0;
    d->x_tap_rs[1] = static_cast<int16_t>(-g->s);
    d->y_ca = g->c_in;
    d->y_atoms = g->s;
    d->y_shift_atoms = 1;
    d->y_taps_mode = 1;
    d->y_groups = (g->r + 1) / 2;
    d->y_ntaps = d->y_groups * g->s;
    d->y_sets = 512 / (g->s * g->c_in) < d->y_groups ? 512 / (g->s * g->c_in) : d->y_groups;  // filter-row pairs per CTA
    if (d->y_sets > 4) d->y_sets = 4;
    for (int q = 0; q < d->y_groups; ++q) {
      const int ry = 2 * q + 1;  // may equal R for odd R: that half is dropped by the tap-id check
      for (int a = 0; a < g->s; ++a) {
        fpg_tap t = fwd_tap(ry, a, 1, g->pad, x->c_stride);
        t.dy -= 1;
        d->y_taps[q * g->s + a] = t;
        d->y_tap_rs[q * g->s + a] = static_cast<int16_t>(ry * g->s + a);
      }
    }
    d->tile_w = 64;
    d->tile_h = 1;
    d->kt_x = ceil_div(wo, 64);
    d->kt_y = ho + 1;
    y_extra = g->s - 1;
    make_act_view(dy, 1, 64, 64, 1, &d->x);
    make_act_view(x, 1, d->y_ca, 64 + y_extra, 1, &d->y);
  } else if (!big_out && big_in && shift_ok && g->c_in == 64 && g->s * g->c_out <= 256 && rows_of_64_ok(wp)) {
    // Few output channels (content / tanh heads): dW[k,c,r,s] = sum_p x[p,c] * dy[p - (r,s) + pad, k] over INPUT
    // pixels p. M = 128 = x at input rows p and p + 1 (filter rows r_y and r_y + 1), N = S x c_out = the S column taps
    // as pixel-shift atoms of one dy box. Input rows run from -1 (see above).
    FPG_REQUIRE(g->c_out == 16 || g->c_out == 32, "unsupported c_out %d", g->c_out);
    d->x_is_dy = 0;
    d->x_ca = 64;
    d->x_atoms = 2;
    d->x_groups = 1;
    d->x_taps_mode = 1;
    d->x_ntaps = 2;
    fpg_tap up = {0, 0, 0, -1};
    d->x_taps[0] = up;
    d->x_taps[1] = null_tap;
    d->x_tap_rs[0] = 0;
    d->x_tap_rs[1] = static_cast<int16_t>(g->s);
    d->y_ca = g->c_out;
    d->y_atoms = g->s;
    d->y_shift_atoms = 1;
    d->y_taps_mode = 1;
    d->y_groups = (g->r + 1) / 2;
    d->y_ntaps = d->y_groups * g->s;
    d->y_sets = 512 / (g->s * g->c_out) < d->y_groups ? 512 / (g->s * g->c_out) : d->y_groups;
    if (d->y_sets > 4) d->y_sets = 4;
    for (int q = 0; q < d->y_groups; ++q) {
      const int ry = 2 * q;
      for (int a = 0; a < g->s; ++a) {
        // atom a reads dy a pixels further right: column tap s = S - 1 - a
        fpg_tap t = {0, g->pad - (g->s - 1) + a, 0, -(ry - g->pad) - 1};
        d->y_taps[q * g->s + a] = t;
        d->y_tap_rs[q * g->s + a] = static_cast<int16_t>(ry * g->s + (g->s - 1 - a));
      }
    }
    d->tile_w = 64;
    d->tile_h = 1;
    d->kt_x = ceil_div(wp, 64);
    d->kt_y = hp + 1;
    y_extra = g->s - 1;
    make_act_view(x, 1, 64, 64, 1, &d->x);
    make_act_view(dy, 1, d->y_ca, 64 + y_extra, 1, &d->y);
  } else if (big_out && big_in && g->c_out == 256 && g->c_in == 256 && ntaps >= 2 && sms >= 2 * ((ntaps + 1) / 2) &&
             getenv("FPG_DISABLE_WGRAD_PAIR") == nullptr) {
    // 256 x 256-channel layers (residual convs): CTA pairs, an item = two taps (igemm_wgrad2_kernel)
    d->cta_pair = 1;
    d->x_is_dy = 1;
    d->x_ca = 64;
    d->x_atoms = 4;
    d->x_groups = 1;
    d->x_taps_mode = 0;
    d->x_ntaps = 1;
    d->x_taps[0] = null_tap;
    d->y_ca = 64;
    d->y_atoms = 4;
    d->y_groups = 1;
    d->y_taps_mode = 0;
    d->y_sets = 2;
    d->y_ntaps = ntaps;
    for (int t = 0; t < ntaps; ++t) {
      d->y_taps[t] = in_taps[t];
      d->y_tap_rs[t] = static_cast<int16_t>(t);
    }
    make_act_view(dy, 1, 64, d->tile_w, d->tile_h, &d->x);
    make_act_view(x, g->stride, 64, d->tile_w, d->tile_h, &d->y);
  } else if (big_out) {
    // X = dy (rows = output channels)
    d->x_is_dy = 1;
    d->tap_on_x = 0;
    d->x_ca = 64;
    d->x_atoms = g->c_out % 256 == 0 && big_in ? 4 : (g->c_out % 128 == 0 ? 2 : 1);
    d->x_groups = g->c_out / (64 * d->x_atoms);
    d->x_taps_mode = 0;
    d->x_ntaps = 1;
    d->x_taps[0] = null_tap;
    make_act_view(dy, 1, 64, d->tile_w, d->tile_h, &d->x);
    if (big_in && g->c_in == 64 && ntaps > 1 && getenv("FPG_DISABLE_WGRAD_TAP_ATOMS") == nullptr) {
      // one 64-channel atom of input: a tap per item would stream dy once per tap for an N = 64 MMA (110 B/clk of
      // operand traffic per SM, 54-cycle MMAs). Up to four taps become atoms of one N <= 256 tile instead.
      d->y_ca = 64;
      d->y_atoms = largest_divisor_le(ntaps, 4);
      d->y_groups = ntaps / d->y_atoms;
      d->y_taps_mode = 1;
    } else if (big_in) {
      d->y_ca = 64;
      int ya = g->c_in / 64;
      if (ya > 4) ya = 4;
      while ((g->c_in / 64) % ya != 0) --ya;
      d->y_atoms = ya;
      d->y_groups = g->c_in / (64 * ya);
      d->y_taps_mode = 0;
    } else {
      FPG_REQUIRE(g->c_in == 16 || g->c_in == 32, "unsupported c_in %d", g->c_in);
      d->y_ca = g->c_in;
      d->y_atoms = largest_divisor_le(ntaps, 256 / g->c_in);
      d->y_groups = ntaps / d->y_atoms;
      d->y_taps_mode = 1;
    }
    d->y_ntaps = ntaps;
    for (int t = 0; t < ntaps; ++t) {
      d->y_taps[t] = in_taps[t];
      d->y_tap_rs[t] = static_cast<int16_t>(t);
    }
    make_act_view(x, g->stride, d->y_ca, d->tile_w, d->tile_h, &d->y);
  } else {
    // small c_out: X = input (rows = input channels), Y = dy
    FPG_REQUIRE(big_in && (g->c_out == 16 || g->c_out == 32), "unsupported wgrad channels c_in %d c_out %d", g->c_in,
                g->c_out);
    d->x_is_dy = 0;
    d->x_ca = 64;
    d->x_atoms = g->c_in % 128 == 0 ? 2 : 1;
    d->x_groups = g->c_in / (64 * d->x_atoms);
    d->x_taps_mode = 0;
    d->y_ca = g->c_out;
    if (g->stride == 1 && ntaps > 1) {
      // Shift dy instead of x: dW[k, (r,s), c] = sum_p x[p, c] * dy[p - (r,s) + pad, k] over INPUT pixels p. Several
      // taps of dy become atoms of one wide N tile, so each x tile is loaded once per filter row instead of once per
      // tap (the tap-per-item form is L2-bandwidth bound: every item streams the whole input again).
      d->tap_on_x = 0;
      d->x_ntaps = 1;
      d->x_taps[0] = null_tap;
      pick_tile(hp, wp, &d->tile_w, &d->tile_h, 64);
      d->kt_x = ceil_div(wp, d->tile_w);
      d->kt_y = ceil_div(hp, d->tile_h);
      d->y_atoms = largest_divisor_le(ntaps, 256 / g->c_out);
      d->y_groups = ntaps / d->y_atoms;
      d->y_taps_mode = 1;
      d->y_ntaps = ntaps;
      for (int r = 0; r < g->r; ++r)
        for (int s = 0; s < g->s; ++s) {
          fpg_tap t = {0, -(s - g->pad), 0, -(r - g->pad)};
          d->y_taps[r * g->s + s] = t;
          d->y_tap_rs[r * g->s + s] = static_cast<int16_t>(r * g->s + s);
        }
    } else {
      d->tap_on_x = 1;
      d->x_ntaps = ntaps;
      for (int t = 0; t < ntaps; ++t) {
        d->x_taps[t] = in_taps[t];
        d->x_tap_rs[t] = static_cast<int16_t>(t);
      }
      d->y_atoms = 1;
      d->y_groups = 1;
      d->y_taps_mode = 0;
      d->y_ntaps = 1;
      d->y_taps[0] = null_tap;
    }
    make_act_view(x, g->stride, 64, d->tile_w, d->tile_h, &d->x);
    make_act_view(dy, 1, d->y_ca, d->tile_w, d->tile_h, &d->y);
  }
  (void)x_extra;
  if (d->y_sets < 1) d->y_sets = 1;
  const int NX = d->x_taps_mode ? d->x_groups : d->x_groups * d->x_ntaps;
  const int sets = d->y_sets;
  const int NY = d->y_taps_mode ? (d->y_groups + sets - 1) / sets : d->y_groups * d->y_ntaps;
  const int items = NX * NY;
  const int total_kt = d->n_img * d->kt_x * d->kt_y;
  int splits = sms / items;
  if (splits < 1) splits = 1;
  if (splits > total_kt) splits = total_kt;
  if (splits > 160) splits = 160;
  d->splits = splits;
  if (d->cta_pair) {
    // (ntaps / 2) tap pairs of `splits` clusters each + (odd tap count) the last tap alone with `last_splits` clusters.
    // Measured per k-tile: 1100 clk for a tap pair (8 MMAs of 128 clk), ~650 clk for the single tap (load bound), so
    // the split counts minimise the longer of the two cluster kinds over the sms / 2 co-resident clusters.
    const int clusters = sms / 2, pairs = ntaps / 2;
    int best_sp = 1, best_last = 1;
    int64_t best_cost = INT64_MAX;
    for (int sp = 1; sp <= clusters && sp <= total_kt; ++sp) {
      int last = (ntaps & 1) ? clusters - pairs * sp : 0;
      if ((ntaps & 1) && last < 1) break;
      if (pairs * sp > clusters) break;
      if (last > total_kt) last = total_kt;
      if (last > sp) last = sp;  // workspace slots are [splits][items]
      const int64_t c2 = static_cast<int64_t>(ceil_div(total_kt, sp)) * 1100;
      const int64_t c1 = (ntaps & 1) ? static_cast<int64_t>(ceil_div(total_kt, last)) * 650 : 0;
      const int64_t cost = c2 > c1 ? c2 : c1;
      if (cost < best_cost) {
        best_cost = cost;
        best_sp = sp;
        best_last = last;
      }
    }
    d->splits = best_sp;
    d->last_splits = (ntaps & 1) ? best_last : 0;
    d->stages = 4;
    return 0;
  }
  const int M = d->x_atoms * d->x_ca;
  const int x_stage = d->x_shift_atoms ? ((64 + d->x_atoms - 1) * d->x_ca * 2 + 1023) / 1024 * 1024 : M * 128;
  const int y_px = 64 + (d->y_shift_atoms ? d->y_atoms - 1 : 0) + (d->y_shifts - 1);
  const int y_stage = ((y_px * d->y_ca * 2 + 1023) / 1024 * 1024) * (d->y_shift_atoms ? d->y_sets : d->y_atoms);
  d->stages = pick_stages(x_stage + y_stage);
  return 0;
}

static int wgrad_items_y(const fpg_igemm_wgrad_desc* d) {
  const int sets = d->y_sets > 1 ? d->y_sets : 1;
  return d->y_taps_mode ? (d->y_groups + sets - 1) / sets : d->y_groups * d->y_ntaps;
}

static int64_t wgrad_ws_floats(const fpg_igemm_wgrad_desc* d) {
  if (d->cta_pair) return static_cast<int64_t>(d->splits) * ((d->y_ntaps + 1) / 2) * 256 * 512;
  const int NX = d->x_taps_mode ? d->x_groups : d->x_groups * d->x_ntaps;
  const int NY = wgrad_items_y(d);
  return static_cast<int64_t>(d->splits) * NX * NY * (d->x_atoms * d->x_ca) * (d->y_atoms * d->y_ca) *
         (d->y_shifts > 1 ? d->y_shifts : 1) * (d->y_sets > 1 ? d->y_sets : 1);
}

// ------------------------------------------------------------------------------------------------ kernels
struct ReduceArgs {
  int32_t x_ca, x_atoms, x_groups, x_taps_mode, x_ntaps;
  int32_t y_ca, y_atoms, y_groups, y_taps_mode, y_ntaps;
  int32_t splits, x_is_dy, y_shifts, y_sets, taps_total, pair, last_splits;
  int32_t parts;  // threads that share one output float4, each summing every parts-th split (power of two <= 16)
  int64_t stride_k, stride_c;
  int32_t k_valid, c_valid;
  int16_t x_tap_rs[FPG_MAX_TAPS];
  int16_t y_tap_rs[FPG_MAX_TAPS];
};

// One thread per 4 consecutive workspace floats (same item, row, tap; 4 consecutive channels of the Y operand): sums the
// split partials with 16-byte loads in a fixed order and scatters into the parameter-layout gradient.
// Small gradients (a few thousand float4 against ~148 splits: PatchGAN model.0, the 1x1 / 7x7 heads, the stem) used to
// run a handful of CTAs with ~148 dependent-latency loop trips each (38-42 us for 64 KB of output): `parts` threads now
// share an output, each summing every parts-th split, combined in shared memory in a fixed order.
template <bool MULTI>  // MULTI: a.parts > 1 (the single-part instantiation keeps the plain loop of the large layers)
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, const ReduceArgs a) {
  __shared__ float4 part_sum[MULTI ? 256 : 1];
  const int parts = MULTI ? a.parts : 1;
  const int M = a.x_atoms * a.x_ca, N = a.y_atoms * a.y_ca;
  const int NT = a.y_shifts * a.y_sets * N;
  const int NX = a.x_taps_mode ? a.x_groups : a.x_groups * a.x_ntaps;
  const int NY = a.pair ? (a.y_ntaps + 1) / 2
                        : (a.y_taps_mode ? (a.y_groups + a.y_sets - 1) / a.y_sets : a.y_groups * a.y_ntaps);
  const int per_split4 = NX * NY * M * (NT >> 2);  // float4 per split (a few million at most)
  const int outs = blockDim.x / parts;              // outputs per block
  const int part = MULTI ? threadIdx.x / outs : 0;
  const int idx4 = blockIdx.x * outs + (MULTI ? threadIdx.x % outs : threadIdx.x);
  bool live = idx4 < per_split4;
  if (!MULTI && !live) return;
  if (!live) {
    part_sum[threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    return;
  }
  // accumulator-native workspace order of an item: [row block of RB rows][16-column chunk][4][RB][4] floats
  const int item4 = M * (NT >> 2);
  const int item = idx4 / item4;
  const int rem = idx4 - item * item4;
  const int RB = M >= 128 ? 32 : 16;
  const int lane = rem % RB, i4 = (rem / RB) & 3, blk = rem / (RB * 4);
  const int nt = (blk % (NT >> 4)) * 16 + i4 * 4;
  const int m = (blk / (NT >> 4)) * RB + lane;
  const int grp = nt / N, n = nt % N;  // MMA group: pixel shift (y_shifts) or tap set (y_sets)
  const int shift = a.y_sets > 1 ? 0 : grp;  // tap sets carry their taps in the table, shift groups add to the id
  const int xi = item / NY, yi = item % NY;
  int xtap, xch, ytap, ych;
  {
    const int atom = m / a.x_ca, within = m % a.x_ca;
    if (a.x_taps_mode) {
      xtap = xi * a.x_atoms + atom;
      xch = within;
    } else {
      xtap = xi / a.x_groups;
      xch = ((xi % a.x_groups) * a.x_atoms + atom) * a.x_ca + within;
    }
  }
  {
    const int atom = n / a.y_ca, within = n % a.y_ca;  // the 4 floats share the atom (y_ca >= 16)
    if (a.y_taps_mode) {
      ytap = (yi * a.y_sets + (a.y_sets > 1 ? grp : 0)) * a.y_atoms + atom;
      ych = within;
    } else if (a.pair) {  // item = tap pair, group = tap inside the pair, atoms = channel chunks
      ytap = yi * 2 + grp;
      ych = atom * a.y_ca + within;
    } else {
      ytap = yi / a.y_groups;
      ych = ((yi % a.y_groups) * a.y_atoms + atom) * a.y_ca + within;
    }
  }
  int tap = -1;
  const int x_valid = a.x_is_dy ? a.k_valid : a.c_valid, y_valid = a.x_is_dy ? a.c_valid : a.k_valid;
  if (xtap < a.x_ntaps && ytap < a.y_ntaps) {  // (else: dummy atoms of a ragged last group)
    tap = a.x_tap_rs[xtap] + a.y_tap_rs[ytap] + shift;
    if (tap >= a.taps_total || xch >= x_valid || ych >= y_valid) tap = -1;
  }
  live = tap >= 0;
  if (!MULTI && !live) return;
  const float4* p = reinterpret_cast<const float4*>(ws) + idx4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int splits = (a.pair && (a.y_ntaps & 1) && yi == NY - 1) ? a.last_splits : a.splits;
  if (live) {
#pragma unroll 8
    for (int s = part; s < splits; s += parts) {  // loads batched by the unroll, summation order fixed
      const float4 v = __ldcg(p + static_cast<int64_t>(s) * per_split4);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
  }
  if (MULTI) {
    part_sum[threadIdx.x] = acc;
    __syncthreads();
    if (part != 0 || !live) return;
    for (int q = 1; q < parts; ++q) {
      const float4 v = part_sum[q * outs + threadIdx.x];
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
  }
  const int64_t x_stride = a.x_is_dy ? a.stride_k : a.stride_c, y_stride = a.x_is_dy ? a.stride_c : a.stride_k;
  float* out = dw + xch * x_stride + ych * y_stride + tap;
  const float r[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (ych + e < y_valid) out[e * y_stride] = r[e];
}

struct PackArgs {
  int32_t rows, taps, cols;      // dst[rows][taps][cols] bf16
  int32_t rows_valid, cols_valid;
  int64_t src_stride_row, src_stride_col;
  int32_t src_tap[FPG_MAX_TAPS];  // source tap index (r*S+s) or -1 for zero
};

__global__ void pack_weights_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, PackArgs a) {
  const int64_t total = static_cast<int64_t>(a.rows) * a.taps * a.cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(i % a.cols);
    const int t = static_cast<int>((i / a.cols) % a.taps);
    const int row = static_cast<int>(i / (static_cast<int64_t>(a.cols) * a.taps));
    float v = 0.f;
    const int st = a.src_tap[t];
    if (st >= 0 && row < a.rows_valid && col < a.cols_valid)
      v = src[row * a.src_stride_row + col * a.src_stride_col + st];
    dst[i] = __float2bfloat16(v);
  }
}

// One launch for every packed operand of a network. Block b works on tile block_first[b] of job block_job[b].
// A weight job gathers dst[row][tap][col] (bf16) = src[row * stride_row + col * stride_col + src_tap[tap]] from the fp32
// parameter [K][C][R*S]: one of (row, col) walks the parameter with stride R*S, the other with stride C*R*S, so per
// element the old kernel fetched a 4-byte value out of a 36-byte neighbourhood that other threads of other blocks wanted
// (116 us for the generator's 24 M packed values). Here a block takes a tile of kPackRows(rs) x 64 (row, col) pairs with
// ALL their taps: it reads the tile as contiguous runs of the parameter into shared memory (the run direction is
// whichever index has stride R*S) and writes every (row, tap) as a contiguous 128-byte run of the destination.
// Jobs with dst_fp32 != 0 copy an fp32 vector (bias) into a zero-padded fp32 buffer, 2048 elements per block.
constexpr int kPackChunk = 2048;
constexpr int kPackCols = 64;
constexpr int kPackUnroll = 8;
constexpr int kPackSmemFloats = 11 * 1024;  // 44 KB (static shared memory: 48 KB with the job copy)
__host__ __device__ inline int pack_rs(const fpg_pack_job& j) {
  return static_cast<int>(j.src_stride_row < j.src_stride_col ? j.src_stride_row : j.src_stride_col);
}
__host__ __device__ inline int pack_tile_rows(int rs) {  // rows per tile such that rows|1 x 64 x (rs|1) floats fit
  int r = kPackSmemFloats / (kPackCols * (rs | 1));
  r = r > 33 ? 32 : r - 1;  // the shared-memory row pitch is (rows | 1)
  return r < 1 ? 1 : r;
}

__global__ void __launch_bounds__(256)
pack_batched_kernel(const fpg_pack_job* __restrict__ jobs, const int32_t* __restrict__ block_job,
                    const int32_t* __restrict__ block_first) {
  __shared__ fpg_pack_job job;
  __shared__ float tile[kPackSmemFloats];
  {
    const int32_t* src = reinterpret_cast<const int32_t*>(jobs + block_job[blockIdx.x]);
    int32_t* dst = reinterpret_cast<int32_t*>(&job);
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(fpg_pack_job) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  if (job.dst_fp32) {
    const int64_t total = static_cast<int64_t>(job.rows) * job.taps * job.cols;
    const int64_t begin = static_cast<int64_t>(block_first[blockIdx.x]) * kPackChunk;
#pragma unroll
    for (int u = 0; u < kPackChunk / 256; ++u) {
      const int64_t i = begin + u * 256 + threadIdx.x;
      if (i >= total) break;
      const int col = static_cast<int>(i % job.cols);
      const int t = static_cast<int>((i / job.cols) % job.taps);
      const int row = static_cast<int>(i / (static_cast<int64_t>(job.cols) * job.taps));
      float v = 0.f;
      const int st = job.src_tap[t];
      if (st >= 0 && row < job.rows_valid && col < job.cols_valid)
        v = job.src[row * job.src_stride_row + col * job.src_stride_col + st];
      static_cast<float*>(job.dst)[i] = v;
    }
    return;
  }
  const int rs = pack_rs(job), rsp = rs | 1;
  const int TR = pack_tile_rows(rs), trp = TR | 1;
  const int col_tiles = (job.cols + kPackCols - 1) / kPackCols;
  const int tile_idx = block_first[blockIdx.x];
  const int row0 = (tile_idx / col_tiles) * TR, col0 = (tile_idx % col_tiles) * kPackCols;
  const int nr = min(TR, job.rows_valid - row0), nc = min(kPackCols, job.cols_valid - col0);  // valid part (may be <= 0)
  const bool col_runs = job.src_stride_col < job.src_stride_row;  // consecutive cols are rs floats apart (fprop layout)
  // ---- load: contiguous runs of the parameter (warp per run, lanes along it; e / rs by reciprocal: e < 64 * 49)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_rs = 1.f / static_cast<float>(rs);
  if (nr > 0 && nc > 0) {
    if (col_runs) {
      const int run = nc * rs;  // per tile row
      for (int r = warp; r < nr; r += 8) {
        const float* sp = job.src + (row0 + r) * job.src_stride_row + static_cast<int64_t>(col0) * rs;
        for (int e0 = lane; e0 < run; e0 += 32 * kPackUnroll) {  // batches of independent loads (latency bound)
          float v[kPackUnroll];
#pragma unroll
          for (int u = 0; u < kPackUnroll; ++u) v[u] = e0 + 32 * u < run ? __ldg(sp + e0 + 32 * u) : 0.f;
#pragma unroll
          for (int u = 0; u < kPackUnroll; ++u) {
            const int e = e0 + 32 * u;
            if (e < run) {
              const int c = __float2int_rz((static_cast<float>(e) + 0.5f) * inv_rs), st = e - c * rs;
              tile[(c * trp + r) * rsp + st] = v[u];
            }
          }
        }
      }
    } else {
      const int run = nr * rs;  // per tile column
      for (int c = warp; c < nc; c += 8) {
        const float* sp = job.src + (col0 + c) * job.src_stride_col + static_cast<int64_t>(row0) * rs;
        for (int e0 = lane; e0 < run; e0 += 32 * kPackUnroll) {
          float v[kPackUnroll];
#pragma unroll
          for (int u = 0; u < kPackUnroll; ++u) v[u] = e0 + 32 * u < run ? __ldg(sp + e0 + 32 * u) : 0.f;
#pragma unroll
          for (int u = 0; u < kPackUnroll; ++u) {
            const int e = e0 + 32 * u;
            if (e < run) {
              const int r = __float2int_rz((static_cast<float>(e) + 0.5f) * inv_rs), st = e - r * rs;
              tile[(c * trp + r) * rsp + st] = v[u];
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- store: dst[row][tap][col], two columns (4 bytes) per thread, 32 lanes = one 128-byte run per (row, tap)
  const int rows_here = min(TR, job.rows - row0), cols_here = min(kPackCols, job.cols - col0);  // cols are even
  const int pairs = cols_here >> 1;
  __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(job.dst);
  for (int rt = warp; rt < rows_here * job.taps; rt += 8) {
    const int r = rt / job.taps, t = rt - r * job.taps;
    const int st = job.src_tap[t];
    const bool live = st >= 0 && r < nr;
    __nv_bfloat16* drow = dst + (static_cast<int64_t>(row0 + r) * job.taps + t) * job.cols + col0;
    for (int cp = lane; cp < pairs; cp += 32) {
      const int c = 2 * cp;
      float v0 = 0.f, v1 = 0.f;
      if (live) {
        if (c < nc) v0 = tile[(c * trp + r) * rsp + st];
        if (c + 1 < nc) v1 = tile[((c + 1) * trp + r) * rsp + st];
      }
      *reinterpret_cast<uint32_t*>(drow + c) = pack_bf16x2(v0, v1);
    }
  }
}

// stats[(i*c + ch)*2] = {mean, rstd} of image i from the epilogue partials [i][rows][c][2] (fixed summation order).
// The kernel is latency bound (1 MB per image): block = 32 channels (16 lanes x 2 channels per 16-byte load) x 64 row
// parts, so that the <= 512 rows of an image are 8 independent loads per thread and a launch has n * c / 32 blocks
// (it was 64 channels x 16 parts: 8.3-9 us per launch, 30 launches per step).
constexpr int kFinalizeParts = 64;
constexpr int kFinalizeLanes = 16;
__global__ void __launch_bounds__(kFinalizeLanes * kFinalizeParts)
stats_finalize_kernel(const float* __restrict__ partial, int rows, int c, float inv_count, float eps, int sums_only,
                      float* __restrict__ stats) {
  const int i = blockIdx.x;
  const int lane = threadIdx.x % kFinalizeLanes, part = threadIdx.x / kFinalizeLanes;
  const int ch = blockIdx.y * (2 * kFinalizeLanes) + 2 * lane;  // c is a multiple of 16: ch and ch + 1 are both valid
  __shared__ float4 red[kFinalizeParts][kFinalizeLanes];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);  // {sum, sumsq} of ch, {sum, sumsq} of ch + 1
  if (ch < c) {
    const float4* p = reinterpret_cast<const float4*>(partial + (static_cast<int64_t>(i) * rows * c + ch) * 2);
    const int64_t row4 = c / 2;  // float4 per partial row
#pragma unroll 8
    for (int r = part; r < rows; r += kFinalizeParts) {
      const float4 v = __ldcg(p + static_cast<int64_t>(r) * row4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  red[part][lane] = acc;
  __syncthreads();
  if (part == 0 && ch < c) {
    for (int q = 1; q < kFinalizeParts; ++q) {
      const float4 v = red[q][lane];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const float s[2] = {acc.x, acc.z}, ss[2] = {acc.y, acc.w};
    float4 out;
    float* o = reinterpret_cast<float*>(&out);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float mean = s[k] * inv_count;
      const float var = fmaxf(ss[k] * inv_count - mean * mean, 0.f);
      o[2 * k] = mean;
      // sums_only: the two plane means themselves (InstanceNorm backward: mean g', mean g' * zhat)
      o[2 * k + 1] = sums_only ? ss[k] * inv_count : rsqrtf(var + eps);
    }
    *reinterpret_cast<float4*>(stats + (static_cast<int64_t>(i) * c + ch) * 2) = out;
  }
}

// partial rows [n][rows][c][2] -> [n][ceil(rows / 256)][c][2] (fixed order): keeps the final reduction short when a
// batch-statistics layer has 10^5 partial rows
__global__ void __launch_bounds__(256)
stats_rows_reduce_kernel(const float* __restrict__ in, int rows, int c, float* __restrict__ out, int out_rows) {
  const int i = blockIdx.x, grp = blockIdx.y;
  const int ch = blockIdx.z * 64 + (threadIdx.x & 63), part = threadIdx.x >> 6;
  __shared__ float2 red[4][64];
  float a = 0.f, b = 0.f;
  if (ch < c) {
    const float2* p = reinterpret_cast<const float2*>(in) + static_cast<int64_t>(i) * rows * c + ch;
    const int r_end = min(rows, (grp + 1) * 256);
#pragma unroll 8
    for (int r = grp * 256 + part; r < r_end; r += 4) {
      const float2 v = __ldcg(p + static_cast<int64_t>(r) * c);
      a += v.x;
      b += v.y;
    }
  }
  red[part][threadIdx.x & 63] = make_float2(a, b);
  __syncthreads();
  if (part == 0 && ch < c) {
    for (int q = 1; q < 4; ++q) {
      a += red[q][threadIdx.x].x;
      b += red[q][threadIdx.x].y;
    }
    reinterpret_cast<float2*>(out)[(static_cast<int64_t>(i) * out_rows + grp) * c + ch] = make_float2(a, b);
  }
}

}  // namespace fpg

using namespace fpg;

extern "C" {

int fpg_conv2d_fprop_plan(const fpg_act* x, const void* w_packed, const float* bias, int act, const fpg_conv_geom* g,
                          const fpg_act* y, int sm_count, fpg_igemm_fprop_desc* out_desc) {
  return plan_fprop(x, w_packed, bias, act, g, y, sm_count, out_desc);
}

int fpg_conv2d_fprop(const fpg_act* x, const void* w_packed, const float* bias, int act, const fpg_conv_geom* g,
                     const fpg_act* y, void* stream) {
  fpg_igemm_fprop_desc d;
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(FPG_ENOTSUP, "no CUDA device");
  fpg_igemm_rows_desc rd;
  int rc = plan_rows(x, w_packed, bias, act, g, y, 0, sms, &rd);
  if (rc == 0) return fpg_igemm_rows_launch(&rd, stream);
  if (rc != 1) return rc;
  rc = plan_fprop(x, w_packed, bias, act, g, y, sms, &d);
  if (rc) return rc;
  return fpg_igemm_fprop_launch(&d, stream);
}

int fpg_conv2d_dgrad_plan(const fpg_act* dy, const void* w_packed_t, const float* bias, int act,
                          const fpg_conv_geom* g, const fpg_act* dx, int sm_count, fpg_igemm_fprop_desc* out_descs,
                          int* n_descs) {
  return plan_dgrad(dy, w_packed_t, bias, act, g, dx, sm_count, out_descs, n_descs);
}

int fpg_conv2d_dgrad(const fpg_act* dy, const void* w_packed_t, const float* bias, int act, const fpg_conv_geom* g,
                     const fpg_act* dx, void* stream) {
  fpg_igemm_fprop_desc d[4];
  int n = 0;
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(FPG_ENOTSUP, "no CUDA device");
  fpg_igemm_rows_desc rd;
  int rc = plan_rows(dy, w_packed_t, bias, act, g, dx, 1, sms, &rd);
  if (rc == 0) return fpg_igemm_rows_launch(&rd, stream);
  if (rc != 1) return rc;
  rc = plan_dgrad(dy, w_packed_t, bias, act, g, dx, sms, d, &n);
  if (rc) return rc;
  rc = fpg_igemm_s2cls_launch(d, n, stream);  // the four parity classes of a stride-2 layer in one launch
  if (rc != 1) return rc;
  for (int q = 0; q < n; ++q) {
    rc = fpg_igemm_fprop_launch(&d[q], stream);
    if (rc) return rc;
  }
  return 0;
}

int32_t fpg_conv2d_dgrad_launches(const fpg_act* dy, const fpg_conv_geom* g, const fpg_act* dx) {
  const int sms = sm_count_cached() > 0 ? sm_count_cached() : 148;
  fpg_igemm_rows_desc rd;
  if (plan_rows(dy, nullptr, nullptr, 0, g, dx, 1, sms, &rd) == 0) return 1;
  fpg_igemm_fprop_desc d[4];
  int n = 0;
  if (plan_dgrad(dy, nullptr, nullptr, 0, g, dx, sms, d, &n)) return -1;
  if (n == 4 && getenv("FPG_DISABLE_S2CLS") == nullptr) {
    // the eligibility test of fpg_igemm_s2cls_launch, without launching
    int slots = 0;
    bool ok = true;
    for (int q = 0; q < 4; ++q) {
      ok = ok && d[q].cblk == 64 && !d[q].cta_pair && d[q].n_blocks == 1 && d[q].block_n <= 64 && d[q].num_taps <= 4 &&
           d[q].tiles_x1 == 0 && d[q].tile_w == d[0].tile_w && d[q].tile_h == d[0].tile_h;
      slots += d[q].num_taps * (d[q].c_per_tap / 64);
    }
    if (ok && static_cast<size_t>(slots) * d[0].block_n * 128 + 2 * 16384 + 2048 <= 227 * 1024) return 1;
  }
  return n;
}

int fpg_conv2d_rows_plan(const fpg_act* a, const void* w_packed, const float* bias, int act, const fpg_conv_geom* g,
                         const fpg_act* out, int dgrad, int sm_count, fpg_igemm_rows_desc* out_desc) {
  return plan_rows(a, w_packed, bias, act, g, out, dgrad, sm_count, out_desc);
}

static int stats_rows_of(const fpg_igemm_fprop_desc* d) {
  return (d->tiles_y * d->tiles_x * (d->cta_pair ? 2 : 1) + d->tiles_y1 * d->tiles_x1) * 4;
}

// rows per image of the epilogue statistics that fpg_conv2d_fprop_stats / _dgrad_stats produce for this layer
// (0: the layer runs on a kernel without that epilogue and the caller uses fpg_instnorm_stats instead)
int32_t fpg_conv_stats_rows(const fpg_act* a, const fpg_conv_geom* g, const fpg_act* out, int dgrad) {
  const int sms = sm_count_cached() > 0 ? sm_count_cached() : 148;
  fpg_igemm_rows_desc rd;
  if (plan_rows(a, nullptr, nullptr, 0, g, out, dgrad, sms, &rd) == 0) return 0;
  fpg_igemm_fprop_desc d[4];
  int n = 1;
  if (dgrad) {
    if (plan_dgrad(a, nullptr, nullptr, 0, g, out, sms, d, &n)) return -1;
  } else {
    if (plan_fprop(a, nullptr, nullptr, 0, g, out, sms, &d[0])) return -1;
  }
  int rows = 0;
  for (int q = 0; q < n; ++q) rows += stats_rows_of(&d[q]);
  return rows;
}

int fpg_conv2d_fprop_stats(const fpg_act* x, const void* w_packed, const float* bias, int act, const fpg_conv_geom* g,
                           const fpg_act* y, float* stat_partial, void* stream) {
  FPG_REQUIRE(stat_partial != nullptr, "null statistics buffer");
  fpg_igemm_fprop_desc d;
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(FPG_ENOTSUP, "no CUDA device");
  fpg_igemm_rows_desc rd;
  FPG_REQUIRE(plan_rows(x, w_packed, bias, act, g, y, 0, sms, &rd) == 1, "layer runs on the row-stationary kernel");
  int rc = plan_fprop(x, w_packed, bias, act, g, y, sms, &d);
  if (rc) return rc;
  d.stat_partial = stat_partial;
  d.stat_rows_per_img = stats_rows_of(&d);
  d.stat_row0 = 0;
  return fpg_igemm_fprop_launch(&d, stream);
}

int fpg_conv2d_dgrad_stats(const fpg_act* dy, const void* w_packed_t, const float* bias, int act,
                           const fpg_conv_geom* g, const fpg_act* dx, float* stat_partial, void* stream) {
  FPG_REQUIRE(stat_partial != nullptr, "null statistics buffer");
  fpg_igemm_fprop_desc d[4];
  int n = 0;
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(FPG_ENOTSUP, "no CUDA device");
  fpg_igemm_rows_desc rd;
  FPG_REQUIRE(plan_rows(dy, w_packed_t, bias, act, g, dx, 1, sms, &rd) == 1, "layer runs on the row-stationary kernel");
  int rc = plan_dgrad(dy, w_packed_t, bias, act, g, dx, sms, d, &n);
  if (rc) return rc;
  int rows = 0;
  for (int q = 0; q < n; ++q) rows += stats_rows_of(&d[q]);
  int row0 = 0;
  for (int q = 0; q < n; ++q) {
    d[q].stat_partial = stat_partial;
    d[q].stat_rows_per_img = rows;
    d[q].stat_row0 = row0;
    row0 += stats_rows_of(&d[q]);
  }
  rc = fpg_igemm_s2cls_launch(d, n, stream);
  if (rc != 1) return rc;
  for (int q = 0; q < n; ++q) {
    rc = fpg_igemm_fprop_launch(&d[q], stream);
    if (rc) return rc;
  }
  return 0;
}

int fpg_conv2d_dgrad_inbwd(const fpg_act* dy, const void* w_packed_t, const fpg_conv_geom* g, const fpg_act* dx,
                           const fpg_act* z, const fpg_act* zprev, const fpg_act* add, float* stat_partial,
                           int32_t* rows_per_img, void* stream) {
  FPG_REQUIRE(dy && w_packed_t && g && dx && z && stat_partial && rows_per_img, "null argument");
  const fpg_act* same[2] = {z, zprev};
  for (const fpg_act* t : same)
    if (t != nullptr)
      FPG_REQUIRE(t->fp32 == FPG_DT_BF16 && t->halo == dx->halo && t->h == dx->h && t->w == dx->w && t->c == dx->c &&
                      t->n == dx->n && t->c % 64 == 0,
                  "z / zprev must be bf16 tensors of dx's geometry (the convolution's saved input and the block input)");
  if (add != nullptr)
    FPG_REQUIRE(add->fp32 == FPG_DT_BF16 && add->h == dx->h && add->w == dx->w && add->c == dx->c && add->n == dx->n,
                "add must match the interior of dx");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(FPG_ENOTSUP, "no CUDA device");
  if (g->stride != 1) return 1;
  fpg_igemm_rows_desc rd;
  if (plan_rows(dy, w_packed_t, nullptr, FPG_ACT_NONE, g, dx, 1, sms, &rd) != 1) return 1;  // row-stationary layer
  fpg_igemm_fprop_desc d[4];
  int n = 0;
  int rc = plan_dgrad(dy, w_packed_t, nullptr, FPG_ACT_NONE, g, dx, sms, d, &n);
  if (rc) return rc;
  if (n != 1 || d[0].cta_pair || d[0].cblk != 64 || d[0].block_n % 64 != 0) return 1;
  // the operand ring needs 3 stages x 8 KB per staged tensor of shared memory: fewer stages for the main loop when
  // it does not fit beside them
  const int n_tensors = 1 + (zprev ? 1 : 0) + (add ? 1 : 0);
  while (d[0].stages > 2 && static_cast<size_t>(d[0].stages) * (16384 + d[0].block_n * 128) +
                                    3 * n_tensors * 8192 + 2048 > 227 * 1024)
    --d[0].stages;
  d[0].stat_partial = stat_partial;
  d[0].stat_rows_per_img = stats_rows_of(&d[0]);
  d[0].stat_row0 = 0;
  d[0].inbwd_mode = zprev ? 2 : 1;
  d[0].inbwd_has_add = add ? 1 : 0;
  d[0].inbwd_h = dx->h;
  d[0].inbwd_w = dx->w;
  d[0].inbwd_halo = dx->halo;
  d[0].inbwd_add_halo = add ? add->halo : 0;
  const bool two = d[0].tiles_x1 > 0 && d[0].tiles_y1 > 0;
  make_act_view(z, 1, 32, d[0].tile_w, d[0].tile_h, &d[0].inbwd_z);
  if (two) make_act_view(z, 1, 32, d[0].tile_w1, d[0].tile_h1, &d[0].inbwd_z1);
  if (zprev) {
    make_act_view(zprev, 1, 32, d[0].tile_w, d[0].tile_h, &d[0].inbwd_prev);
    if (two) make_act_view(zprev, 1, 32, d[0].tile_w1, d[0].tile_h1, &d[0].inbwd_prev1);
  }
  if (add) {
    make_act_view(add, 1, 32, d[0].tile_w, d[0].tile_h, &d[0].inbwd_add);
    if (two) make_act_view(add, 1, 32, d[0].tile_w1, d[0].tile_h1, &d[0].inbwd_add1);
  }
  *rows_per_img = d[0].stat_rows_per_img;
  return fpg_igemm_fprop_launch(&d[0], stream);
}

static int reduce_stat_rows(const float* stat_partial, int32_t rows_per_img, int32_t n, int32_t c, float inv_count,
                            float eps, int sums_only, float* out, void* stream) {
  const float* cur = stat_partial;
  int rows = rows_per_img;
  // long row lists are first folded 256:1 into the tail of the buffer (the caller sizes it for that)
  float* spare = const_cast<float*>(stat_partial) + static_cast<int64_t>(n) * rows_per_img * c * 2;
  while (rows > 512) {
    const int out_rows = (rows + 255) / 256;
    stats_rows_reduce_kernel<<<dim3(n, out_rows, (c + 63) / 64), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        cur, rows, c, spare, out_rows);
    cur = spare;
    spare += static_cast<int64_t>(n) * out_rows * c * 2;
    rows = out_rows;
  }
  if (c % 2 != 0 || (reinterpret_cast<uintptr_t>(cur) & 15) != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0)
    return fail(FPG_EINVAL, "statistics finalize: channels must be even and the buffers 16-byte aligned");
  stats_finalize_kernel<<<dim3(n, (c + 2 * kFinalizeLanes - 1) / (2 * kFinalizeLanes)),
                          kFinalizeLanes * kFinalizeParts, 0, static_cast<cudaStream_t>(stream)>>>(
      cur, rows, c, inv_count, eps, sums_only, out);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int fpg_instnorm_bwd_sums_finalize(const float* stat_partial, int32_t rows_per_img, int32_t n, int32_t c,
                                   int64_t count_per_img, float* red, void* stream) {
  FPG_REQUIRE(stat_partial && red && rows_per_img > 0 && n > 0 && c > 0 && count_per_img > 0, "bad argument");
  return reduce_stat_rows(stat_partial, rows_per_img, n, c, 1.f / static_cast<float>(count_per_img), 0.f, 1, red, stream);
}

int fpg_instnorm_stats_finalize(const float* stat_partial, int32_t rows_per_img, int32_t n, int32_t c,
                                int64_t count_per_img, float eps, float* stats, void* stream) {
  FPG_REQUIRE(stat_partial && stats && rows_per_img > 0 && n > 0 && c > 0 && count_per_img > 0, "bad argument");
  return reduce_stat_rows(stat_partial, rows_per_img, n, c, 1.f / static_cast<float>(count_per_img), eps, 0, stats, stream);
}

int fpg_conv2d_wgrad_plan(const fpg_act* x, const fpg_act* dy, const fpg_conv_geom* g, int sm_count,
                          fpg_igemm_wgrad_desc* out_desc) {
  return plan_wgrad(x, dy, g, sm_count, out_desc);
}

int64_t fpg_conv2d_wgrad_ws_bytes(const fpg_act* x, const fpg_act* dy, const fpg_conv_geom* g, int sm_count) {
  fpg_igemm_wgrad_desc d;
  if (plan_wgrad(x, dy, g, sm_count, &d)) return -1;
  return wgrad_ws_floats(&d) * 4;
}

int fpg_conv2d_wgrad(const fpg_act* x, const fpg_act* dy, const fpg_conv_geom* g, float* dw, int64_t dw_stride_k,
                     int64_t dw_stride_c, int32_t k_valid, int32_t c_valid, float* ws, void* stream) {
  fpg_igemm_wgrad_desc d;
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(FPG_ENOTSUP, "no CUDA device");
  int rc = plan_wgrad(x, dy, g, sms, &d);
  if (rc) return rc;
  d.ws = ws;
  rc = fpg_igemm_wgrad_launch(&d, stream);
  if (rc) return rc;
  ReduceArgs a;
  a.x_ca = d.x_ca;
  a.x_atoms = d.x_atoms;
  a.x_groups = d.x_groups;
  a.x_taps_mode = d.x_taps_mode;
  a.x_ntaps = d.x_ntaps;
  a.y_ca = d.y_ca;
  a.y_atoms = d.y_atoms;
  a.y_groups = d.y_groups;
  a.y_taps_mode = d.y_taps_mode;
  a.y_ntaps = d.y_ntaps;
  a.splits = d.splits;
  a.x_is_dy = d.x_is_dy;
  a.y_shifts = d.y_shifts > 1 ? d.y_shifts : 1;
  a.y_sets = d.y_sets > 1 ? d.y_sets : 1;
  a.pair = d.cta_pair ? 1 : 0;
  a.last_splits = d.last_splits;
  a.taps_total = d.taps_r * d.taps_s;
  a.stride_k = dw_stride_k;
  a.stride_c = dw_stride_c;
  a.k_valid = k_valid;
  a.c_valid = c_valid;
  for (int i = 0; i < FPG_MAX_TAPS; ++i) {
    a.x_tap_rs[i] = d.x_tap_rs[i];
    a.y_tap_rs[i] = d.y_tap_rs[i];
  }
  const int64_t per_split4 = wgrad_ws_floats(&d) / d.splits / 4;
  const int threads = 256;
  // fewer outputs than ~2 CTAs per SM: share every output among `parts` threads (see the kernel)
  int parts = 1;
  while (parts < 16 && parts * 2 <= d.splits && (per_split4 * parts + threads - 1) / threads < 2 * sms) parts *= 2;
  a.parts = parts;
  const int64_t blocks = (per_split4 * parts + threads - 1) / threads;
  if (parts > 1)
    wgrad_reduce_kernel<true><<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(ws, dw, a);
  else
    wgrad_reduce_kernel<false><<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(ws, dw, a);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int64_t fpg_packed_weight_bytes(const fpg_conv_geom* g) {
  const int cblk = pick_cblk(g->c_in);
  if (cblk < 0) return -1;
  return static_cast<int64_t>(g->c_out) * padded_taps(g->r * g->s, g->c_in, cblk) * g->c_in * 2;
}

int64_t fpg_packed_weight_dgrad_bytes(const fpg_conv_geom* g) {
  if (pick_cblk(g->c_out) < 0) return -1;
  int64_t k_of[4], off_of[4];
  int nc = 0;
  dgrad_class_layout(g, k_of, off_of, &nc);
  return (off_of[nc - 1] + k_of[nc - 1] * g->c_in) * 2;
}

int fpg_dgrad_class_info(const fpg_conv_geom* g, int cls, int32_t* src_tap, int32_t* num_taps_padded,
                         int64_t* elem_offset, int32_t* num_classes) {
  FPG_REQUIRE(g && src_tap && num_taps_padded && elem_offset && num_classes, "null argument");
  FPG_REQUIRE(pick_cblk(g->c_out) > 0, "unsupported c_out %d", g->c_out);
  int64_t k_of[4], off_of[4];
  int nc = 0;
  dgrad_class_layout(g, k_of, off_of, &nc);
  *num_classes = nc;
  FPG_REQUIRE(cls >= 0 && cls < nc, "class %d of %d", cls, nc);
  int rs[FPG_MAX_TAPS];
  fpg_tap taps[FPG_MAX_TAPS];
  const int nt = dgrad_class_taps(g, cls >> 1, cls & 1, rs, taps);
  *num_taps_padded = static_cast<int32_t>(k_of[cls] / g->c_out);
  *elem_offset = off_of[cls];
  for (int t = 0; t < FPG_MAX_TAPS; ++t) src_tap[t] = t < nt ? rs[t] : -1;
  return 0;
}

int fpg_pack_weights(const float* src, int64_t src_stride_k, int64_t src_stride_c, int32_t k_valid, int32_t c_valid,
                     const fpg_conv_geom* g, void* dst, void* stream) {
  const int cblk = pick_cblk(g->c_in);
  FPG_REQUIRE(cblk > 0, "unsupported c_in %d", g->c_in);
  PackArgs a;
  a.rows = g->c_out;
  a.taps = padded_taps(g->r * g->s, g->c_in, cblk);
  a.cols = g->c_in;
  a.rows_valid = k_valid;
  a.cols_valid = c_valid;
  a.src_stride_row = src_stride_k;
  a.src_stride_col = src_stride_c;
  for (int t = 0; t < FPG_MAX_TAPS; ++t) a.src_tap[t] = t < g->r * g->s ? t : -1;
  const int64_t total = static_cast<int64_t>(a.rows) * a.taps * a.cols;
  const int threads = 256;
  int64_t blocks = (total + threads - 1) / threads;
  if (blocks > 4096) blocks = 4096;
  pack_weights_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), a);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int fpg_pack_weights_dgrad(const float* src, int64_t src_stride_k, int64_t src_stride_c, int32_t k_valid,
                           int32_t c_valid, const fpg_conv_geom* g, void* dst, void* stream) {
  const int cblk = pick_cblk(g->c_out);
  FPG_REQUIRE(cblk > 0, "unsupported c_out %d", g->c_out);
  int64_t k_of[4], off_of[4];
  int nc = 0;
  dgrad_class_layout(g, k_of, off_of, &nc);
  for (int q = 0; q < nc; ++q) {
    int rs[FPG_MAX_TAPS];
    fpg_tap taps[FPG_MAX_TAPS];
    const int nt = dgrad_class_taps(g, q >> 1, q & 1, rs, taps);
    PackArgs a;
    a.rows = g->c_in;  // GEMM N side = forward input channels
    a.taps = static_cast<int32_t>(k_of[q] / g->c_out);
    a.cols = g->c_out;
    a.rows_valid = c_valid;
    a.cols_valid = k_valid;
    a.src_stride_row = src_stride_c;
    a.src_stride_col = src_stride_k;
    for (int t = 0; t < FPG_MAX_TAPS; ++t) a.src_tap[t] = t < nt ? rs[t] : -1;
    const int64_t total = static_cast<int64_t>(a.rows) * a.taps * a.cols;
    const int threads = 256;
    int64_t blocks = (total + threads - 1) / threads;
    if (blocks > 4096) blocks = 4096;
    pack_weights_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        src, static_cast<__nv_bfloat16*>(dst) + off_of[q], a);
    FPG_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}

int fpg_pack_jobs(const float* src, int64_t src_stride_k, int64_t src_stride_c, int32_t k_valid, int32_t c_valid,
                  const fpg_conv_geom* g, void* dst_fprop, void* dst_dgrad, fpg_pack_job* jobs, int32_t* n_jobs) {
  FPG_REQUIRE(src && g && jobs && n_jobs, "null argument");
  int n = 0;
  if (dst_fprop != nullptr) {
    const int cblk = pick_cblk(g->c_in);
    FPG_REQUIRE(cblk > 0, "unsupported c_in %d", g->c_in);
    fpg_pack_job* j = &jobs[n++];
    memset(j, 0, sizeof(*j));
    j->src = src;
    j->dst = dst_fprop;
    j->rows = g->c_out;
    j->taps = padded_taps(g->r * g->s, g->c_in, cblk);
    j->cols = g->c_in;
    j->rows_valid = k_valid;
    j->cols_valid = c_valid;
    j->src_stride_row = src_stride_k;
    j->src_stride_col = src_stride_c;
    for (int t = 0; t < FPG_MAX_TAPS; ++t) j->src_tap[t] = static_cast<int8_t>(t < g->r * g->s ? t : -1);
  }
  if (dst_dgrad != nullptr) {
    const int cblk = pick_cblk(g->c_out);
    FPG_REQUIRE(cblk > 0, "unsupported c_out %d", g->c_out);
    int64_t k_of[4], off_of[4];
    int nc = 0;
    dgrad_class_layout(g, k_of, off_of, &nc);
    for (int q = 0; q < nc; ++q) {
      int rs[FPG_MAX_TAPS];
      fpg_tap taps[FPG_MAX_TAPS];
      const int nt = dgrad_class_taps(g, q >> 1, q & 1, rs, taps);
      fpg_pack_job* j = &jobs[n++];
      memset(j, 0, sizeof(*j));
      j->src = src;
      j->dst = static_cast<__nv_bfloat16*>(dst_dgrad) + off_of[q];
      j->rows = g->c_in;
      j->taps = static_cast<int32_t>(k_of[q] / g->c_out);
      j->cols = g->c_out;
      j->rows_valid = c_valid;
      j->cols_valid = k_valid;
      j->src_stride_row = src_stride_c;
      j->src_stride_col = src_stride_k;
      for (int t = 0; t < FPG_MAX_TAPS; ++t) j->src_tap[t] = static_cast<int8_t>(t < nt ? rs[t] : -1);
    }
  }
  *n_jobs = n;
  return 0;
}

int fpg_pack_job_copy_f32(const float* src, int32_t count_valid, float* dst, int32_t count_padded, fpg_pack_job* job) {
  FPG_REQUIRE(src && dst && job && count_valid <= count_padded, "bad argument");
  memset(job, 0, sizeof(*job));
  job->src = src;
  job->dst = dst;
  job->rows = 1;
  job->taps = 1;
  job->cols = count_padded;
  job->rows_valid = 1;
  job->cols_valid = count_valid;
  job->src_stride_row = 0;
  job->src_stride_col = 1;
  job->dst_fp32 = 1;
  for (int t = 0; t < FPG_MAX_TAPS; ++t) job->src_tap[t] = static_cast<int8_t>(t == 0 ? 0 : -1);
  return 0;
}

int32_t fpg_pack_job_blocks(const fpg_pack_job* job) {
  if (job->dst_fp32) {
    const int64_t total = static_cast<int64_t>(job->rows) * job->taps * job->cols;
    return static_cast<int32_t>((total + kPackChunk - 1) / kPackChunk);
  }
  const int tr = pack_tile_rows(pack_rs(*job));
  return ((job->rows + tr - 1) / tr) * ((job->cols + kPackCols - 1) / kPackCols);
}

int fpg_pack_weights_batched(const fpg_pack_job* jobs_dev, const int32_t* block_job_dev, const int32_t* block_first_dev,
                             int32_t n_blocks, void* stream) {
  FPG_REQUIRE(jobs_dev && block_job_dev && block_first_dev && n_blocks > 0, "bad argument");
  pack_batched_kernel<<<n_blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(jobs_dev, block_job_dev,
                                                                               block_first_dev);
  FPG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"
