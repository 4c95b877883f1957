/*
 * oracle/raster.c - CPU oracle for the shape -> pixels draw loop.   TEST INFRASTRUCTURE ONLY.
 *
 * Restates, as scalar C, what the reference does between "compiled paths" and "pixels":
 *   ts/src/lib/renderers/canvas-renderer.ts:69-78    renderStage   (clear to transparent, scale 1/20)
 *   ts/src/lib/renderers/canvas-renderer.ts:114-129  drawShape     (CTM = scale(1/20) . matrix, paths in order)
 *   ts/src/lib/renderers/canvas-renderer.ts:190-267  drawMorphShape/drawMorphPath (lerp end*r + start*(1-r))
 *   ts/src/lib/renderers/canvas-renderer.ts:269-350  drawPath      (beginPath .. fill(): nonzero, implicit close,
 *                                                    per-path source-over; Solid / Bitmap pattern / Focal gradient)
 *   ts/src/lib/css-color.ts:11-13                    fromNormalizedColor
 * and the part of the un-vendored native backend those calls land in (npm canvas@2.6.1 -> Cairo -> pixman,
 * ts/package.json:39): curve flattening, anti-aliased non-zero fill, pattern / gradient sampling and the
 * premultiplied 8-bit OVER operator.  Cairo's own scan converter is not in the reference tree and is NOT
 * reproduced sample-for-sample; this oracle defines:
 *   - geometry in 24.8 fixed point (as Cairo), curves flattened by uniform subdivision at 0.1 px,
 *   - exact-area coverage accumulated in Q16 integers per 16x16 tile from tile-clipped edge records,
 *   - pixman's 8-bit arithmetic for IN/OVER (MUL_UN8 with rounding),
 *   - box-footprint ("GOOD") bitmap filtering, 256-interval gradient ramps.
 * Every arithmetic step is IEEE-754 single/double without contraction (build with -ffp-contract=off) or
 * integer, so the CUDA path can and must match it bit for bit (edges, bin counts AND pixels).
 *
 * Pinned against the reference's PNG goldens by tests/test_oracle_golden.py (tolerances there).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SWFO_TILE 16
#define SWFO_TILE_FX 4096     /* 16 px in 1/256 px units */
#define SWFO_MAXLEN_FX 16384  /* flattened edges are at most 64 px long in x and in y */
#define SWFO_CLAMP_PX 32768.0 /* device coordinates are clamped to +-32768 px before 24.8 conversion */

enum { SWFO_PAINT_SOLID = 0, SWFO_PAINT_LINEAR = 1, SWFO_PAINT_FOCAL = 2, SWFO_PAINT_BITMAP = 3 };
enum { SWFO_SPREAD_PAD = 0, SWFO_SPREAD_REFLECT = 1, SWFO_SPREAD_REPEAT = 2 };

typedef struct {
  double s[6];      /* start state x0,y0,cx,cy,x1,y1 in twips (cx,cy ignored for lines) */
  double e[6];      /* end state (morph); equals s for static shapes */
  int32_t is_curve; /* 0 line, 1 quadratic */
  int32_t path;     /* path index local to the definition */
} swfo_segment;

typedef struct {
  int32_t type;       /* SWFO_PAINT_* */
  int32_t spread;     /* SWFO_SPREAD_* (gradients) */
  int32_t repeating;  /* bitmap */
  int32_t bitmap;     /* index into scene bitmaps, -1 if none */
  uint8_t color0[4];  /* straight RGBA8 (solid; morph start) */
  uint8_t color1[4];  /* straight RGBA8 morph end (== color0 for static) */
  int32_t color_is_morph; /* colour goes through the morph lerp + css-color path (canvas-renderer.ts:241-250) */
  double matrix[6];   /* fill matrix scale_x, rotate_skew0, rotate_skew1, scale_y, tx, ty (fill space -> twips) */
  double focal;       /* focal point in [-1,1] */
  const uint32_t *lut; /* SWFO_RAMP_SIZE premultiplied RGBA8 entries, entry k = the gradient at (k + 1/2) / size */
  int32_t sampled;    /* != 0: stroke outline - coverage by 15 sub-scanlines per pixel row with the non-zero rule applied
                         per sub-scanline (sampled_coverage), because the outline overlaps itself */
} swfo_paint;

typedef struct {
  int32_t first_seg, n_seg;   /* into scene segs */
  int32_t first_path, n_path; /* into scene paints */
  int32_t is_morph;
} swfo_def;

typedef struct {
  int32_t def;
  float m[6];     /* Matrix2D order: scale_x, scale_y, rotate_skew0, rotate_skew1, tx, ty  (rs/src/stage.rs:12-20) */
  uint16_t ratio; /* MorphRatio (rs/src/stage.rs:28-34) */
  uint16_t use_ratio_f; /* != 0: ratio_f (the TypeScript renderer's number in 0..1, display/morph-shape.ts) replaces ratio */
  float ratio_f;
  int32_t has_cx; /* != 0: cx is a swf-tree ColorTransformWithAlpha for this draw (include/swfr.h swfr_color_transform) */
  int16_t cx[8];  /* red, green, blue, alpha mult (Sfixed8P8 epsilons), then red, green, blue, alpha add */
} swfo_item;

typedef struct {
  int32_t w, h;
  const uint8_t *rgba; /* premultiplied RGBA8, row-major */
} swfo_bitmap;

typedef struct {
  int32_t width, height;
  int32_t n_items;
  const swfo_item *items;
  const swfo_def *defs;
  const swfo_segment *segs;
  const swfo_paint *paints;
  const swfo_bitmap *bitmaps;
  uint32_t background; /* premultiplied RGBA8 every pixel starts from: 0 = clearRect (canvas-renderer.ts:70-72, the
                          default), or the opaque stage colour like the windowed Rust renderer (gfx_renderer.rs:292-301) */
} swfo_scene;

/* optional debug taps (may be NULL) */
typedef struct {
  int32_t *edges;        /* x0,y0,x1,y1 per edge (24.8) */
  int32_t *edge_path;    /* path-instance index per edge */
  int64_t edges_cap;     /* capacity in edges */
  int64_t n_edges;       /* out */
  uint32_t *tile_counts; /* tiles_x*tiles_y, += binned records per tile */
  int64_t n_records;     /* out: total binned records */
  int64_t n_slots_drawn; /* out: (path,tile) pairs composited */
} swfo_debug;

/* ------------------------------------------------------------------------------------------ */
/* integer helpers                                                                            */
/* ------------------------------------------------------------------------------------------ */

static inline int64_t floordiv64(int64_t a, int64_t b) { /* b > 0 */
  int64_t q = a / b;
  if ((a % b != 0) && (a < 0)) q -= 1;
  return q;
}
/* round-half-up of a/b, b != 0 */
static inline int64_t rdiv64(int64_t a, int64_t b) {
  if (b < 0) {
    a = -a;
    b = -b;
  }
  return floordiv64(2 * a + b, 2 * b);
}
static inline int32_t imin32(int32_t a, int32_t b) { return a < b ? a : b; }
static inline int32_t imax32(int32_t a, int32_t b) { return a > b ? a : b; }
static inline int32_t iabs32(int32_t a) { return a < 0 ? -a : a; }

/* ------------------------------------------------------------------------------------------ */
/* stage 1: lerp + transform + flatten  (canvas-renderer.ts:24-26, 179-188, 212-239, 274-290)  */
/* ------------------------------------------------------------------------------------------ */

static inline double lerp_ref(double start, double end, double r) { /* canvas-renderer.ts:24-26 */
  double a = end * r;
  double b = 1.0 - r;
  double c = start * b;
  return a + c;
}

/* twips -> device px -> 24.8 fixed.  m = the canvas CTM in Matrix2D order: the reference starts every frame with
 * ctx.scale(1 / 20, 1 / 20) (canvas-renderer.ts:74) and multiplies the shape matrix in with ctx.transform
 * (canvas-renderer.ts:179-188), i.e. Cairo holds CTM = 0.05 * M entry by entry (ctm_from_matrix below) and transforms
 * path points with it at path-build time: x' = xx x + xy y + x0. */
static inline void ctm_from_matrix(const double m[6], double ctm[6]) {
  for (int i = 0; i < 6; i++) ctm[i] = m[i] * 0.05;
}
static inline void to_device_fx(const double m[6], double x, double y, int32_t *fx, int32_t *fy) {
  double px = m[0] * x;
  double qx = m[3] * y;
  double sx = px + qx;
  sx = sx + m[4];
  double py = m[2] * x;
  double qy = m[1] * y;
  double sy = py + qy;
  sy = sy + m[5];
  if (!(sx > -SWFO_CLAMP_PX)) sx = -SWFO_CLAMP_PX; /* also catches NaN */
  if (sx > SWFO_CLAMP_PX) sx = SWFO_CLAMP_PX;
  if (!(sy > -SWFO_CLAMP_PX)) sy = -SWFO_CLAMP_PX;
  if (sy > SWFO_CLAMP_PX) sy = SWFO_CLAMP_PX;
  *fx = (int32_t)llrint(sx * 256.0);
  *fy = (int32_t)llrint(sy * 256.0);
}

/* number of line pieces for a segment given its fixed-point control points */
static int32_t piece_count(int is_curve, const int32_t p[6]) {
  if (!is_curve) {
    int32_t ext = imax32(iabs32(p[4] - p[0]), iabs32(p[5] - p[1]));
    int32_t n = (ext + SWFO_MAXLEN_FX - 1) / SWFO_MAXLEN_FX;
    return n < 1 ? 1 : n;
  }
  int64_t ddx = (int64_t)p[0] - 2 * (int64_t)p[2] + p[4];
  int64_t ddy = (int64_t)p[1] - 2 * (int64_t)p[3] + p[5];
  int64_t m2 = ddx * ddx + ddy * ddy;
  /* smallest n >= 1 with 262144 n^4 >= 25 m2   <=>   |dd| / (4 n^2) <= 0.1 px */
  int64_t rhs = 25 * m2;
  int64_t n = (int64_t)sqrt(sqrt((double)m2 * (25.0 / 262144.0)));
  if (n < 1) n = 1;
  if (n > 2048) n = 2048;
  while (n < 2048 && 262144 * n * n * n * n < rhs) n++;
  while (n > 1 && 262144 * (n - 1) * (n - 1) * (n - 1) * (n - 1) >= rhs) n--;
  int32_t leg = imax32(imax32(iabs32(p[2] - p[0]), iabs32(p[3] - p[1])), imax32(iabs32(p[4] - p[2]), iabs32(p[5] - p[3])));
  int64_t nlen = (2 * (int64_t)leg + SWFO_MAXLEN_FX - 1) / SWFO_MAXLEN_FX;
  if (nlen > n) n = nlen;
  if (n > 4096) n = 4096;
  return (int32_t)n;
}

static inline void piece_point(int is_curve, const int32_t p[6], int32_t n, int32_t i, int32_t *x, int32_t *y) {
  if (!is_curve) {
    *x = p[0] + (int32_t)rdiv64(((int64_t)p[4] - p[0]) * i, n);
    *y = p[1] + (int32_t)rdiv64(((int64_t)p[5] - p[1]) * i, n);
  } else {
    int64_t a = (int64_t)(n - i) * (n - i), b = 2 * (int64_t)i * (n - i), c = (int64_t)i * i, nn = (int64_t)n * n;
    *x = (int32_t)rdiv64(a * p[0] + b * p[2] + c * p[4], nn);
    *y = (int32_t)rdiv64(a * p[1] + b * p[3] + c * p[5], nn);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* stage 2: tile binning                                                                      */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  uint16_t xa, ya, xb, yb; /* tile-relative 1/256 px, 0..4096 */
  uint8_t flag_s, flag_e;  /* start / end was clipped at the tile's left boundary */
} record_t;

typedef struct {
  /* per path-instance tile grid (bbox clipped to the viewport) */
  int32_t bx0, by0, bw, bh;
  int32_t *count;    /* bw*bh */
  int32_t *backdrop; /* bw*bh, deltas then prefix-summed */
  int32_t *offset;   /* bw*bh + 1 */
  int32_t *cursor;
  record_t *records;
} grid_t;

typedef void (*emit_fn)(void *ctx, int32_t tx, int32_t ty, const record_t *r);
typedef void (*backdrop_fn)(void *ctx, int32_t tx_first, int32_t ty, int32_t delta);

static inline int32_t xat(int32_t x0, int32_t y0, int32_t x1, int32_t y1, int32_t Y) {
  return x0 + (int32_t)rdiv64(((int64_t)Y - y0) * ((int64_t)x1 - x0), (int64_t)y1 - y0);
}

/* Walk one flattened edge over the tile grid: emits tile-clipped records and backdrop deltas. */
static void bin_edge(int32_t x0, int32_t y0, int32_t x1, int32_t y1, int32_t tiles_x, int32_t tiles_y, void *ctx,
                     emit_fn emit, backdrop_fn backdrop) {
  const int32_t B = SWFO_TILE_FX;
  int32_t b_first, b_last;
  if (y0 == y1) {
    b_first = b_last = y0 >> 12;
  } else {
    int32_t ylo = imin32(y0, y1), yhi = imax32(y0, y1);
    b_first = ylo >> 12;
    b_last = (yhi - 1) >> 12;
  }
  if (b_first < 0) b_first = 0;
  if (b_last > tiles_y - 1) b_last = tiles_y - 1;
  for (int32_t b = b_first; b <= b_last; b++) {
    int32_t Yt = b * B, Yb = Yt + B;
    int32_t xs, ys, xe, ye;
    if (y0 == y1) {
      xs = x0, ys = y0, xe = x1, ye = y1;
    } else if (y0 < y1) {
      ys = imax32(y0, Yt);
      ye = imin32(y1, Yb);
      xs = (ys == y0) ? x0 : xat(x0, y0, x1, y1, ys);
      xe = (ye == y1) ? x1 : xat(x0, y0, x1, y1, ye);
    } else {
      ys = imin32(y0, Yb);
      ye = imax32(y1, Yt);
      xs = (ys == y0) ? x0 : xat(x0, y0, x1, y1, ys);
      xe = (ye == y1) ? x1 : xat(x0, y0, x1, y1, ye);
    }
    /* winding entering / leaving through the band's top line counts for every tile to the right */
    if (ys == Yt) backdrop(ctx, (xs >> 12) + 1, b, +1);
    if (ye == Yt) backdrop(ctx, (xe >> 12) + 1, b, -1);
    int32_t c0 = imin32(xs, xe) >> 12, c1 = imax32(xs, xe) >> 12;
    if (c0 < 0) c0 = 0;
    if (c1 > tiles_x - 1) c1 = tiles_x - 1;
    for (int32_t t = c0; t <= c1; t++) {
      int32_t X0 = t * B, X1 = X0 + B;
      int32_t ax, ay, bx, by;
      record_t r;
      r.flag_s = r.flag_e = 0;
      if (xs < X0) {
        ax = X0, ay = ys + (int32_t)rdiv64(((int64_t)X0 - xs) * ((int64_t)ye - ys), (int64_t)xe - xs), r.flag_s = 1;
      } else if (xs > X1) {
        ax = X1, ay = ys + (int32_t)rdiv64(((int64_t)X1 - xs) * ((int64_t)ye - ys), (int64_t)xe - xs);
      } else {
        ax = xs, ay = ys;
      }
      if (xe < X0) {
        bx = X0, by = ys + (int32_t)rdiv64(((int64_t)X0 - xs) * ((int64_t)ye - ys), (int64_t)xe - xs), r.flag_e = 1;
      } else if (xe > X1) {
        bx = X1, by = ys + (int32_t)rdiv64(((int64_t)X1 - xs) * ((int64_t)ye - ys), (int64_t)xe - xs);
      } else {
        bx = xe, by = ye;
      }
      if (ay == by && !r.flag_s && !r.flag_e) continue; /* contributes nothing */
      r.xa = (uint16_t)(ax - X0);
      r.ya = (uint16_t)(ay - Yt);
      r.xb = (uint16_t)(bx - X0);
      r.yb = (uint16_t)(by - Yt);
      emit(ctx, t, b, &r);
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* stage 3: per-tile coverage (Q16 signed area, non-zero via min(|acc|,1))                    */
/* ------------------------------------------------------------------------------------------ */

static void accumulate_record(const record_t *rc, int32_t acc[16][16]) {
  int32_t xa = rc->xa, ya = rc->ya, xb = rc->xb, yb = rc->yb;
  /* left-boundary crossing terms: H(y)[row] = clamp((row+1)*256 - y, 0, 256) * 256 */
  if (rc->flag_s || rc->flag_e) {
    int32_t yc = rc->flag_s ? ya : yb;
    int32_t sgn = rc->flag_s ? -1 : +1;
    for (int r = 0; r < 16; r++) {
      int32_t hq = (r + 1) * 256 - yc;
      hq = hq < 0 ? 0 : (hq > 256 ? 256 : hq);
      if (hq)
        for (int i = 0; i < 16; i++) acc[r][i] += sgn * hq * 256;
    }
  }
  if (ya == yb) return;
  const float k = 1.0f / 256.0f;
  float xaf = (float)xa * k, yaf = (float)ya * k, xbf = (float)xb * k, ybf = (float)yb * k;
  float slope = (xbf - xaf) / (ybf - yaf);
  float xlo = xaf < xbf ? xaf : xbf, xhi = xaf < xbf ? xbf : xaf;
  int32_t s = yb > ya ? 1 : -1;
  int32_t ylo = imin32(ya, yb), yhi = imax32(ya, yb);
  for (int r = ylo >> 8; r <= ((yhi - 1) >> 8); r++) {
    int32_t yt = imax32(ylo, 256 * r), ybm = imin32(yhi, 256 * (r + 1));
    int32_t D = s * (ybm - yt) * 256;
    float Df = (float)D;
    float ytf = (float)yt * k, ybmf = (float)ybm * k;
    float xt = fmaf(ytf - yaf, slope, xaf), xm = fmaf(ybmf - yaf, slope, xaf);
    xt = fminf(fmaxf(xt, xlo), xhi);
    xm = fminf(fmaxf(xm, xlo), xhi);
    float xmin = fminf(xt, xm), xmax = fmaxf(xt, xm);
    float w = xmax - xmin;
    float inv2w = w > 0.0f ? 0.5f / w : 0.0f;
    for (int i = 0; i < 16; i++) {
      float fi = (float)i, fi1 = (float)(i + 1);
      int32_t c;
      if (fi1 <= xmin) {
        c = 0;
      } else if (fi >= xmax) {
        c = D;
      } else {
        float u0 = fmaxf(fi - xmin, 0.0f);
        float u1 = fminf(fi1 - xmin, w);
        float a0 = (u0 * u0) * inv2w;
        float a1 = fmaf(u1 * u1, inv2w, fmaxf(fi1 - xmax, 0.0f));
        float f = a1 - a0;
        f = fminf(fmaxf(f, 0.0f), 1.0f);
        c = (int32_t)lrintf(Df * f);
      }
      acc[r][i] += c;
    }
  }
}

/* Coverage of a path that may overlap itself (stroke outlines: joins, caps and inner loops are separate pieces of
 * one contour).  The signed-area integral clamped per pixel over-covers where two parts of the path overlap inside a
 * partly covered pixel; Cairo's scan converter (cairo-tor-scan-converter.c, GRID_Y = 15, GRID_X = 256) applies the
 * fill rule per sub-scanline instead, and so does this: every pixel row is sampled on 15 lines y = row + k / 15 (the
 * top of each sub-row, like tor), a record covers the sample lines in [y_lo, y_hi) of its own extent (half open, so
 * the pieces of one edge in neighbouring tiles never count a line twice), the crossings of a line are walked from
 * left to right with the non-zero rule, and each covered span adds its exact horizontal extent (1/256 px) to the pixels
 * it touches.  mask = coverage in 1/(15 * 256) of a pixel -> 8 bits with tor's GRID_AREA_TO_ALPHA for a 2*256*15 grid. */
#define SWFO_SUBROWS 15
#define SWFO_MAX_SAMPLED 96 /* more records than this in one slot: the area integral is used (a blob of tiny edges) */

static void sampled_coverage(const record_t *rec, int32_t nrec, int32_t backdrop, uint32_t mask[16][16]) {
  int32_t cov[16][18];
  memset(cov, 0, sizeof cov);
  for (int k = 0; k < 16 * SWFO_SUBROWS; k++) {
    const int32_t Yk = k * 256 + 128; /* sample line (centre of the sub-row) in units of 1/(256 * 15) px */
    const int row = k / SWFO_SUBROWS;
    int32_t w = backdrop;
    for (int32_t i = 0; i < nrec; i++) { /* winding at the tile's left edge on this line */
      const record_t *r = &rec[i];
      if (r->flag_s || r->flag_e) {
        int32_t yc15 = (r->flag_s ? r->ya : r->yb) * SWFO_SUBROWS;
        if (Yk >= yc15) w += r->flag_s ? -1 : 1;
      }
    }
    if (w != 0) cov[row][0] += 256;
    int32_t last_x = -1, last_i = -1;
    for (;;) { /* the crossings of this line in ascending (x, record index) order */
      int32_t best_x = INT32_MAX, best_i = -1;
      for (int32_t i = 0; i < nrec; i++) {
        const record_t *r = &rec[i];
        if (r->ya == r->yb) continue;
        int32_t lo15 = imin32(r->ya, r->yb) * SWFO_SUBROWS, hi15 = imax32(r->ya, r->yb) * SWFO_SUBROWS;
        if (Yk < lo15 || Yk >= hi15) continue;
        float slope = (float)(r->xb - r->xa) / (float)((r->yb - r->ya) * SWFO_SUBROWS);
        float x = fmaf((float)(Yk - r->ya * SWFO_SUBROWS), slope, (float)r->xa);
        float xlo = (float)imin32(r->xa, r->xb), xhi = (float)imax32(r->xa, r->xb);
        x = fminf(fmaxf(x, xlo), xhi);
        int32_t xi = (int32_t)lrintf(x);
        if (!(xi > last_x || (xi == last_x && i > last_i))) continue;
        if (xi < best_x) best_x = xi, best_i = i;
      }
      if (best_i < 0) break;
      int32_t s = rec[best_i].yb > rec[best_i].ya ? 1 : -1;
      int32_t w2 = w + s;
      int32_t e = (w == 0 && w2 != 0) ? 1 : ((w != 0 && w2 == 0) ? -1 : 0);
      if (e) {
        int32_t p = best_x >> 8, f = best_x & 255;
        cov[row][p] += e * (256 - f);
        cov[row][p + 1] += e * f;
      }
      w = w2;
      last_x = best_x, last_i = best_i;
    }
  }
  for (int r = 0; r < 16; r++) {
    int32_t run = 0;
    for (int i = 0; i < 16; i++) {
      run += cov[r][i];
      int32_t c = run < 0 ? 0 : (run > 256 * SWFO_SUBROWS ? 256 * SWFO_SUBROWS : run);
      mask[r][i] = (uint32_t)(34 * c + 256) >> 9;
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* stage 4: paint evaluation + pixman-style OVER                                              */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  int32_t type, spread, repeating;
  uint32_t solid;       /* premultiplied, R | G<<8 | B<<16 | A<<24 */
  float inv[6];         /* device px -> fill space: ia, ib, ic, id, itx, ity   (gx = ia*X + ic*Y + itx) */
  float focal, omf;     /* focal point, 1 - focal^2 */
  float inv_omf;        /* 1 / omf (float division) */
  float rx, ry;         /* bitmap footprint in texels */
  const uint32_t *lut;
  const swfo_bitmap *bmp;
  int valid;
  const int16_t *cx; /* colour transform of the draw (NULL: none) */
} paint_inst;

static inline uint32_t mul_un8(uint32_t a, uint32_t b) { /* pixman MUL_UN8 */
  uint32_t t = a * b + 0x80;
  return (t + (t >> 8)) >> 8;
}
static inline uint32_t mul_un8x4(uint32_t p, uint32_t m) {
  return mul_un8(p & 255, m) | (mul_un8((p >> 8) & 255, m) << 8) | (mul_un8((p >> 16) & 255, m) << 16) |
         (mul_un8(p >> 24, m) << 24);
}
/* dst = (src IN m) OVER dst, 8-bit premultiplied */
static inline uint32_t over_masked(uint32_t dst, uint32_t src, uint32_t m) {
  if (m == 0) return dst;
  uint32_t s = (m == 255) ? src : mul_un8x4(src, m);
  uint32_t sa = s >> 24;
  if (sa == 255) return s;
  if (sa == 0 && s == 0) return dst;
  return s + mul_un8x4(dst, 255 - sa); /* cannot overflow for valid premultiplied input */
}

/* Cairo colour -> 16-bit premultiplied -> top 8 bits (pixman solid image) */
static uint32_t solid_premul8(double r8, double g8, double b8, double alpha) {
  double af = (double)(float)alpha; /* node-canvas parses the CSS alpha as float32 */
  if (af < 0) af = 0;
  if (af > 1) af = 1;
  uint32_t a16 = (uint32_t)(af * 65535.0 + 0.5);
  uint32_t r16 = (uint32_t)(((r8 / 255.0) * af) * 65535.0 + 0.5);
  uint32_t g16 = (uint32_t)(((g8 / 255.0) * af) * 65535.0 + 0.5);
  uint32_t b16 = (uint32_t)(((b8 / 255.0) * af) * 65535.0 + 0.5);
  return (r16 >> 8) | ((g16 >> 8) << 8) | ((b16 >> 8) << 16) | ((a16 >> 8) << 24);
}

/* swf-tree ColorTransformWithAlpha on one straight 8-bit channel (the reference renderer has no colour-transform input;
 * these are the SWF semantics, defined by this file: "parity unpinned") */
static int cx_channel(int c, int mult, int add) {
  int v = (int)(((int64_t)c * mult) >> 8) + add; /* floor, also for negative products */
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}
/* ... on a premultiplied pixel (gradients, bitmaps): un-premultiply, transform, premultiply - integers throughout */
static uint32_t cx_premul(uint32_t p, const int16_t *cx) {
  int a = (int)(p >> 24);
  int a2 = cx_channel(a, cx[3], cx[7]);
  uint32_t out = (uint32_t)a2 << 24;
  for (int k = 0; k < 3; k++) {
    int c = (int)((p >> (8 * k)) & 255);
    int s = a ? (c * 255 + a / 2) / a : 0;
    if (s > 255) s = 255;
    int c2 = cx_channel(s, cx[k], cx[4 + k]);
    out |= (uint32_t)((c2 * a2 + 127) / 255) << (8 * k);
  }
  return out;
}
/* ... on a solid fill: the colour is transformed before it is premultiplied; alpha goes through 8 bits */
static uint32_t cx_solid(double r8, double g8, double b8, double alpha, const int16_t *cx) {
  double af = (double)(float)alpha;
  if (af < 0) af = 0;
  if (af > 1) af = 1;
  int a8 = (int)(af * 255.0 + 0.5);
  int r2 = cx_channel((int)r8, cx[0], cx[4]), g2 = cx_channel((int)g8, cx[1], cx[5]), b2 = cx_channel((int)b8, cx[2], cx[6]);
  int a2 = cx_channel(a8, cx[3], cx[7]);
  return solid_premul8((double)r2, (double)g2, (double)b2, a2 / 255.0);
}

/* css-color.ts:11-13 through node-canvas' rgba() parser, for a lerped normalised colour */
static uint32_t morph_solid(const uint8_t c0[4], const uint8_t c1[4], double r, const int16_t *cx) {
  double ch[4];
  for (int i = 0; i < 4; i++) ch[i] = lerp_ref(c0[i] / 255.0, c1[i] / 255.0, r);
  /* red: (r * 0xff) & 0xff  (ToInt32 truncation) */
  double red = (double)(((int64_t)(ch[0] * 255.0)) & 0xff);
  /* green/blue: printed as floats, parsed by node-canvas (float32) and rounded up to an integer channel */
  double g = ceil((double)(float)(ch[1] * 255.0));
  double b = ceil((double)(float)(ch[2] * 255.0));
  if (g < 0) g = 0;
  if (g > 255) g = 255;
  if (b < 0) b = 0;
  if (b > 255) b = 255;
  return cx ? cx_solid(red, g, b, ch[3], cx) : solid_premul8(red, g, b, ch[3]);
}

static void make_paint(const swfo_scene *sc, const swfo_paint *p, const double m[6], double ratio, const int16_t *cx,
                       paint_inst *out) {
  int is_morph = p->color_is_morph;
  memset(out, 0, sizeof(*out));
  out->type = p->type;
  out->spread = p->spread;
  out->repeating = p->repeating;
  out->valid = 1;
  out->cx = cx;
  if (p->type == SWFO_PAINT_SOLID) {
    if (is_morph)
      out->solid = morph_solid(p->color0, p->color1, ratio, cx);
    else if (cx)
      out->solid = cx_solid(p->color0[0], p->color0[1], p->color0[2], p->color0[3] / 255.0, cx);
    else
      out->solid = solid_premul8(p->color0[0], p->color0[1], p->color0[2], p->color0[3] / 255.0);
    return;
  }
  /* combined matrix C = M o F (fill space -> twips), then /20 -> device px.  canvas-renderer.ts:313,321 */
  const double *f = p->matrix; /* scale_x, rotate_skew0, rotate_skew1, scale_y, tx, ty */
  double ma = m[0], md = m[1], mb = m[2], mc = m[3], mtx = m[4], mty = m[5];
  double fa = f[0], fb = f[1], fc = f[2], fd = f[3], ftx = f[4], fty = f[5];
  double ca = (ma * fa + mc * fb) / 20.0;
  double cb = (mb * fa + md * fb) / 20.0;
  double cc = (ma * fc + mc * fd) / 20.0;
  double cd = (mb * fc + md * fd) / 20.0;
  double ctx = ((ma * ftx + mc * fty) + mtx) / 20.0;
  double cty = ((mb * ftx + md * fty) + mty) / 20.0;
  double det = ca * cd - cb * cc;
  if (!(det != 0.0) || !isfinite(det)) {
    out->valid = 0;
    return;
  }
  double ia = cd / det, ib = -cb / det, ic = -cc / det, id = ca / det;
  double itx = -(ia * ctx + ic * cty), ity = -(ib * ctx + id * cty);
  out->inv[0] = (float)ia;
  out->inv[1] = (float)ib;
  out->inv[2] = (float)ic;
  out->inv[3] = (float)id;
  out->inv[4] = (float)itx;
  out->inv[5] = (float)ity;
  if (p->type == SWFO_PAINT_BITMAP) {
    if (p->bitmap < 0) {
      out->valid = 0;
      return;
    }
    out->bmp = &sc->bitmaps[p->bitmap];
    /* Cairo CAIRO_FILTER_GOOD: footprint = |d(source)/d(device)|, bilinear below 4/3, capped at 16 */
    double dx = sqrt(ia * ia + ic * ic), dy = sqrt(ib * ib + id * id);
    if (dx > 16.0) dx = 16.0;
    if (dy > 16.0) dy = 16.0;
    if (dx < 1.0 / 0.75) dx = 1.0;
    if (dy < 1.0 / 0.75) dy = 1.0;
    out->rx = (float)dx;
    out->ry = (float)dy;
  } else {
    double fp = p->type == SWFO_PAINT_FOCAL ? p->focal : 0.0;
    if (fp > 0.98) fp = 0.98;
    if (fp < -0.98) fp = -0.98;
    out->focal = (float)fp;
    out->omf = (float)(1.0 - fp * fp);
    out->inv_omf = 1.0f / out->omf;
    out->lut = p->lut;
  }
}

static inline int32_t floormod(int32_t a, int32_t n) {
  int32_t r = a % n;
  return r < 0 ? r + n : r;
}

/* Every a*b+c below is a fused multiply-add (C fmaf is correctly rounded, the kernel's fmaf() compiles to FFMA): the
 * same operations in the same order on both sides. */
#define SWFO_RAMP_SIZE 1024

static uint32_t eval_paint(const paint_inst *p, int X, int Y) {
  if (p->type == SWFO_PAINT_SOLID) return p->solid;
  float xc = (float)X + 0.5f, yc = (float)Y + 0.5f;
  float gx = fmaf(p->inv[0], xc, fmaf(p->inv[2], yc, p->inv[4]));
  float gy = fmaf(p->inv[1], xc, fmaf(p->inv[3], yc, p->inv[5]));
  if (p->type == SWFO_PAINT_BITMAP) {
    const swfo_bitmap *bm = p->bmp;
    float hrx = p->rx * 0.5f, hry = p->ry * 0.5f;
    float irx = 1.0f / p->rx, iry = 1.0f / p->ry;
    float lox = gx - hrx, hix = gx + hrx, loy = gy - hry, hiy = gy + hry;
    int i0 = (int)floorf(lox), j0 = (int)floorf(loy);
    float acc[4] = {0, 0, 0, 0};
    for (int j = j0; (float)j < hiy; j++) {
      float wl = fmaxf(loy, (float)j), wh = fminf(hiy, (float)(j + 1));
      float wy = fmaxf(wh - wl, 0.0f) * iry;
      int jj = j;
      if (p->repeating)
        jj = floormod(j, bm->h);
      else if (j < 0 || j >= bm->h)
        continue;
      for (int i = i0; (float)i < hix; i++) {
        float vl = fmaxf(lox, (float)i), vh = fminf(hix, (float)(i + 1));
        float wx = fmaxf(vh - vl, 0.0f) * irx;
        int ii = i;
        if (p->repeating)
          ii = floormod(i, bm->w);
        else if (i < 0 || i >= bm->w)
          continue;
        const uint8_t *t = bm->rgba + 4 * ((size_t)jj * bm->w + ii);
        float wgt = wx * wy;
        acc[0] = fmaf(wgt, (float)t[0], acc[0]);
        acc[1] = fmaf(wgt, (float)t[1], acc[1]);
        acc[2] = fmaf(wgt, (float)t[2], acc[2]);
        acc[3] = fmaf(wgt, (float)t[3], acc[3]);
      }
    }
    uint32_t o = 0;
    for (int c = 0; c < 4; c++) {
      float v = fminf(fmaxf(acc[c], 0.0f), 255.0f);
      o |= (uint32_t)lrintf(v) << (8 * c);
    }
    return o;
  }
  /* gradients: canvas-renderer.ts:320-331 (focal/radial, GRAD_RADIUS 16384); linear per SWF gradient square */
  float t;
  if (p->type == SWFO_PAINT_LINEAR) {
    t = fmaf(gx, 1.0f / 32768.0f, 0.5f);
  } else {
    float nx = gx * (1.0f / 16384.0f), ny = gy * (1.0f / 16384.0f);
    float dx = nx - p->focal;
    float c = p->omf * (ny * ny);
    float disc = fmaf(dx, dx, c);
    float s = sqrtf(disc);
    float num = fmaf(p->focal, dx, s);
    t = num * p->inv_omf;
  }
  if (p->spread == SWFO_SPREAD_PAD) {
    t = fminf(fmaxf(t, 0.0f), 1.0f);
  } else if (p->spread == SWFO_SPREAD_REPEAT) {
    t = t - floorf(t);
    t = fminf(fmaxf(t, 0.0f), 1.0f);
  } else {
    float u = t * 0.5f;
    u = u - floorf(u);
    u = u * 2.0f;
    t = u > 1.0f ? 2.0f - u : u;
    t = fminf(fmaxf(t, 0.0f), 1.0f);
  }
  int i = (int)floorf(t * (float)SWFO_RAMP_SIZE);
  if (i < 0) i = 0;
  if (i > SWFO_RAMP_SIZE - 1) i = SWFO_RAMP_SIZE - 1;
  return p->lut[i];
}

/* ------------------------------------------------------------------------------------------ */
/* per path-instance driver                                                                   */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  grid_t *g;
  int pass;
  swfo_debug *dbg;
  int32_t tiles_x;
} bin_ctx;

static void cb_emit(void *vctx, int32_t tx, int32_t ty, const record_t *r) {
  bin_ctx *c = (bin_ctx *)vctx;
  grid_t *g = c->g;
  int32_t lx = tx - g->bx0, ly = ty - g->by0;
  if (lx < 0 || lx >= g->bw || ly < 0 || ly >= g->bh) abort(); /* bbox must cover every record */
  int32_t slot = ly * g->bw + lx;
  if (c->pass == 0) {
    g->count[slot]++;
    if (c->dbg && c->dbg->tile_counts) c->dbg->tile_counts[ty * c->tiles_x + tx]++;
    if (c->dbg) c->dbg->n_records++;
  } else {
    g->records[g->offset[slot] + g->cursor[slot]++] = *r;
  }
}
static void cb_backdrop(void *vctx, int32_t tx_first, int32_t ty, int32_t delta) {
  bin_ctx *c = (bin_ctx *)vctx;
  grid_t *g = c->g;
  if (c->pass != 0) return;
  int32_t ly = ty - g->by0;
  if (ly < 0 || ly >= g->bh) return;
  int32_t lx = tx_first - g->bx0;
  if (lx < 0) lx = 0;
  if (lx >= g->bw) return;
  g->backdrop[ly * g->bw + lx] += delta;
}

int swfo_render(const swfo_scene *sc, uint8_t *out_premul_rgba, swfo_debug *dbg) {
  const int W = sc->width, H = sc->height;
  const int tiles_x = (W + SWFO_TILE - 1) / SWFO_TILE, tiles_y = (H + SWFO_TILE - 1) / SWFO_TILE;
  uint32_t *fb = (uint32_t *)calloc((size_t)W * H, 4); /* clearRect => transparent black (canvas-renderer.ts:70-71) */
  if (!fb) return -1;
  if (sc->background)
    for (size_t i = 0; i < (size_t)W * H; i++) fb[i] = sc->background;
  if (dbg) {
    dbg->n_edges = 0;
    dbg->n_records = 0;
    dbg->n_slots_drawn = 0;
    if (dbg->tile_counts) memset(dbg->tile_counts, 0, sizeof(uint32_t) * tiles_x * tiles_y);
  }
  int32_t *ebuf = NULL;
  size_t ecap = 0;
  int64_t path_inst = 0;
  for (int it = 0; it < sc->n_items; it++) {
    const swfo_item *item = &sc->items[it];
    const swfo_def *def = &sc->defs[item->def];
    double m0[6], m[6];
    for (int i = 0; i < 6; i++) m0[i] = (double)item->m[i];
    ctm_from_matrix(m0, m);
    double ratio = item->use_ratio_f ? (double)item->ratio_f : (double)item->ratio / 65535.0;
    /* an identity transform is no transform (un-premultiplying and premultiplying a pixel again is lossy) */
    static const int16_t cx_identity[8] = {256, 256, 256, 256, 0, 0, 0, 0};
    const int16_t *item_cx = (item->has_cx && memcmp(item->cx, cx_identity, sizeof cx_identity) != 0) ? item->cx : NULL;
    for (int lp = 0; lp < def->n_path; lp++, path_inst++) {
      /* ---- flatten this path's segments ---- */
      size_t ne = 0;
      int32_t minx = INT32_MAX, miny = INT32_MAX, maxx = INT32_MIN, maxy = INT32_MIN;
      for (int si = 0; si < def->n_seg; si++) {
        const swfo_segment *sg = &sc->segs[def->first_seg + si];
        if (sg->path != lp) continue;
        double c[6];
        for (int i = 0; i < 6; i++) c[i] = def->is_morph ? lerp_ref(sg->s[i], sg->e[i], ratio) : sg->s[i];
        int32_t p[6];
        to_device_fx(m, c[0], c[1], &p[0], &p[1]);
        to_device_fx(m, c[4], c[5], &p[4], &p[5]);
        if (sg->is_curve)
          to_device_fx(m, c[2], c[3], &p[2], &p[3]);
        else
          p[2] = p[0], p[3] = p[1];
        int32_t n = piece_count(sg->is_curve, p);
        if ((ne + n) * 4 > ecap) {
          ecap = (ne + n) * 8 + 1024;
          ebuf = (int32_t *)realloc(ebuf, ecap * sizeof(int32_t));
        }
        int32_t px, py;
        piece_point(sg->is_curve, p, n, 0, &px, &py);
        for (int i = 1; i <= n; i++) {
          int32_t qx, qy;
          piece_point(sg->is_curve, p, n, i, &qx, &qy);
          int32_t *e = ebuf + 4 * ne++;
          e[0] = px, e[1] = py, e[2] = qx, e[3] = qy;
          px = qx, py = qy;
        }
        /* bbox over the control polygon (convex hull bounds the curve) */
        for (int k = 0; k < 6; k += 2) {
          minx = imin32(minx, p[k]);
          maxx = imax32(maxx, p[k]);
          miny = imin32(miny, p[k + 1]);
          maxy = imax32(maxy, p[k + 1]);
        }
      }
      if (dbg && dbg->edges) {
        for (size_t k = 0; k < ne; k++) {
          if (dbg->n_edges < dbg->edges_cap) {
            memcpy(dbg->edges + 4 * dbg->n_edges, ebuf + 4 * k, 16);
            if (dbg->edge_path) dbg->edge_path[dbg->n_edges] = (int32_t)path_inst;
          }
          dbg->n_edges++;
        }
      } else if (dbg) {
        dbg->n_edges += (int64_t)ne;
      }
      if (ne == 0) continue;
      /* ---- paint ---- */
      paint_inst paint;
      make_paint(sc, &sc->paints[def->first_path + lp], m0, ratio, item_cx, &paint);
      /* A path whose paint cannot be evaluated, or whose solid colour is premultiplied 0 (alpha 0: over() returns the
       * destination unchanged), composites nothing: it is flattened (the edge tap above lists it) but not binned. */
      if (!paint.valid || (paint.type == SWFO_PAINT_SOLID && paint.solid == 0)) continue;
      /* ---- tile grid over the bbox ---- */
      grid_t g;
      int32_t bx0 = minx >> 12, bx1 = maxx >> 12, by0 = miny >> 12, by1 = maxy >> 12;
      if (bx0 < 0) bx0 = 0;
      if (by0 < 0) by0 = 0;
      if (bx1 > tiles_x - 1) bx1 = tiles_x - 1;
      if (by1 > tiles_y - 1) by1 = tiles_y - 1;
      if (bx1 < bx0 || by1 < by0) continue;
      g.bx0 = bx0, g.by0 = by0, g.bw = bx1 - bx0 + 1, g.bh = by1 - by0 + 1;
      int32_t ns = g.bw * g.bh;
      g.count = (int32_t *)calloc(ns, 4);
      g.backdrop = (int32_t *)calloc(ns, 4);
      g.offset = (int32_t *)calloc(ns + 1, 4);
      g.cursor = (int32_t *)calloc(ns, 4);
      bin_ctx bc = {&g, 0, dbg, tiles_x};
      for (size_t k = 0; k < ne; k++) {
        int32_t *e = ebuf + 4 * k;
        bin_edge(e[0], e[1], e[2], e[3], tiles_x, tiles_y, &bc, cb_emit, cb_backdrop);
      }
      for (int32_t i = 0; i < ns; i++) g.offset[i + 1] = g.offset[i] + g.count[i];
      g.records = (record_t *)malloc(sizeof(record_t) * (size_t)(g.offset[ns] + 1));
      bc.pass = 1;
      for (size_t k = 0; k < ne; k++) {
        int32_t *e = ebuf + 4 * k;
        bin_edge(e[0], e[1], e[2], e[3], tiles_x, tiles_y, &bc, cb_emit, cb_backdrop);
      }
      for (int32_t ly = 0; ly < g.bh; ly++) /* prefix sum of backdrop deltas along each tile row */
        for (int32_t lx = 1; lx < g.bw; lx++) g.backdrop[ly * g.bw + lx] += g.backdrop[ly * g.bw + lx - 1];
      {
        for (int32_t ly = 0; ly < g.bh; ly++) {
          for (int32_t lx = 0; lx < g.bw; lx++) {
            int32_t slot = ly * g.bw + lx;
            if (g.count[slot] == 0 && g.backdrop[slot] == 0) continue;
            if (dbg) dbg->n_slots_drawn++;
            int32_t acc[16][16];
            uint32_t smask[16][16];
            const int32_t nrec = g.offset[slot + 1] - g.offset[slot];
            const int sampled = sc->paints[def->first_path + lp].sampled && nrec > 0 && nrec <= SWFO_MAX_SAMPLED;
            if (sampled) {
              sampled_coverage(&g.records[g.offset[slot]], nrec, g.backdrop[slot], smask);
            } else {
              for (int r = 0; r < 16; r++)
                for (int i = 0; i < 16; i++) acc[r][i] = g.backdrop[slot] * 65536;
              for (int32_t k = g.offset[slot]; k < g.offset[slot + 1]; k++) accumulate_record(&g.records[k], acc);
            }
            int X0 = (g.bx0 + lx) * SWFO_TILE, Y0 = (g.by0 + ly) * SWFO_TILE;
            for (int r = 0; r < 16; r++) {
              int Y = Y0 + r;
              if (Y >= H) break;
              for (int i = 0; i < 16; i++) {
                int X = X0 + i;
                if (X >= W) break;
                uint32_t mcov;
                if (sampled) {
                  mcov = smask[r][i];
                } else {
                  int32_t a = acc[r][i];
                  if (a < 0) a = -a;
                  if (a > 65536) a = 65536;
                  mcov = ((uint32_t)a * 255u + 32768u) >> 16;
                }
                if (!mcov) continue;
                uint32_t src = eval_paint(&paint, X, Y);
                if (paint.cx && paint.type != SWFO_PAINT_SOLID) src = cx_premul(src, paint.cx);
                fb[(size_t)Y * W + X] = over_masked(fb[(size_t)Y * W + X], src, mcov);
              }
            }
          }
        }
      }
      free(g.count);
      free(g.backdrop);
      free(g.offset);
      free(g.cursor);
      free(g.records);
    }
  }
  memcpy(out_premul_rgba, fb, (size_t)W * H * 4);
  free(fb);
  free(ebuf);
  return 0;
}

/* PNG-export semantics (node-canvas toBuffer("image/png") -> Cairo unpremultiply_data) */
void swfo_unpremultiply(const uint8_t *src, uint8_t *dst, int64_t n_px) {
  for (int64_t i = 0; i < n_px; i++) {
    uint32_t a = src[4 * i + 3];
    if (a == 0) {
      dst[4 * i] = dst[4 * i + 1] = dst[4 * i + 2] = dst[4 * i + 3] = 0;
    } else {
      for (int c = 0; c < 3; c++) dst[4 * i + c] = (uint8_t)((src[4 * i + c] * 255u + a / 2) / a);
      dst[4 * i + 3] = (uint8_t)a;
    }
  }
}

/* bitmap upload semantics: straight RGBA8 -> premultiplied RGBA8 */
void swfo_premultiply(const uint8_t *src, uint8_t *dst, int64_t n_px) {
  for (int64_t i = 0; i < n_px; i++) {
    uint32_t a = src[4 * i + 3];
    for (int c = 0; c < 3; c++) dst[4 * i + c] = (uint8_t)((src[4 * i + c] * a + 127u) / 255u);
    dst[4 * i + 3] = (uint8_t)a;
  }
}
