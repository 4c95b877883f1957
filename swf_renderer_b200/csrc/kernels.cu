// CUDA kernels of libswfr_b200 (sm_100a): flatten -> tile binning -> per-tile coverage + paint + blend.
//
// Arithmetic contract: every step below is either integer or IEEE-754 without contraction (this file is
// compiled with -fmad=false, default -prec-div/-prec-sqrt) and follows the same operation order as the CPU
// oracle, so edges, bin counts and pixels are bit-identical to it.  Reference semantics being implemented:
//   ts/src/lib/renderers/canvas-renderer.ts:24-26    lerp(start, end, ratio) = end*r + start*(1-r)
//   ts/src/lib/renderers/canvas-renderer.ts:69-78    clear to transparent, CTM starts as scale(1/20)
//   ts/src/lib/renderers/canvas-renderer.ts:179-188  applyMatrix -> x' = sx*x + rs1*y + tx, y' = rs0*x + sy*y + ty
//   ts/src/lib/renderers/canvas-renderer.ts:269-350  one path = one fill(): non-zero, source-over, in order
//   ts/src/lib/css-color.ts:11-13                    colour channel quantisation of morph-lerped colours
#include <cstdio>

#include "kernels.h"

namespace swfr {
namespace {

// ======================================================================================================
// integer helpers (same definitions as the oracle)
// ======================================================================================================

__device__ __forceinline__ long long floordiv64(long long a, long long b) {  // b > 0
  long long q = a / b;
  if ((a % b != 0) && (a < 0)) q -= 1;
  return q;
}
__device__ __forceinline__ long long rdiv64(long long a, long long b) {  // round half up, b != 0
  if (b < 0) {
    a = -a;
    b = -b;
  }
  return floordiv64(2 * a + b, 2 * b);
}
__device__ __forceinline__ int iabs32(int a) { return a < 0 ? -a : a; }

// ======================================================================================================
// K1: morph lerp + CTM + quadratic flattening
// ======================================================================================================

__device__ __forceinline__ double lerp_ref(double start, double end, double r) {
  double a = end * r;
  double b = 1.0 - r;
  double c = start * b;
  return a + c;
}

__device__ __forceinline__ void to_device_fx(const float *m, double x, double y, int &fx, int &fy) {
  double px = (double)m[0] * x;
  double qx = (double)m[3] * y;
  double sx = px + qx;
  sx = sx + (double)m[4];
  sx = sx / 20.0;
  double py = (double)m[2] * x;
  double qy = (double)m[1] * y;
  double sy = py + qy;
  sy = sy + (double)m[5];
  sy = sy / 20.0;
  if (!(sx > -32768.0)) sx = -32768.0;
  if (sx > 32768.0) sx = 32768.0;
  if (!(sy > -32768.0)) sy = -32768.0;
  if (sy > 32768.0) sy = 32768.0;
  fx = (int)__double2ll_rn(sx * 256.0);
  fy = (int)__double2ll_rn(sy * 256.0);
}

__device__ int piece_count(bool curve, const int *p) {
  if (!curve) {
    int ext = max(iabs32(p[4] - p[0]), iabs32(p[5] - p[1]));
    int n = (ext + kMaxLenFx - 1) / kMaxLenFx;
    return n < 1 ? 1 : n;
  }
  long long ddx = (long long)p[0] - 2 * (long long)p[2] + p[4];
  long long ddy = (long long)p[1] - 2 * (long long)p[3] + p[5];
  long long m2 = ddx * ddx + ddy * ddy;
  long long rhs = 25 * m2;
  long long n = (long long)sqrt(sqrt((double)m2 * (25.0 / 262144.0)));
  if (n < 1) n = 1;
  if (n > 2048) n = 2048;
  while (n < 2048 && 262144 * n * n * n * n < rhs) n++;
  while (n > 1 && 262144 * (n - 1) * (n - 1) * (n - 1) * (n - 1) >= rhs) n--;
  int leg = max(max(iabs32(p[2] - p[0]), iabs32(p[3] - p[1])), max(iabs32(p[4] - p[2]), iabs32(p[5] - p[3])));
  long long nlen = (2 * (long long)leg + kMaxLenFx - 1) / kMaxLenFx;
  if (nlen > n) n = nlen;
  if (n > 4096) n = 4096;
  return (int)n;
}

__device__ __forceinline__ void piece_point(bool curve, const int *p, int n, int i, int &x, int &y) {
  if (!curve) {
    x = p[0] + (int)rdiv64(((long long)p[4] - p[0]) * i, n);
    y = p[1] + (int)rdiv64(((long long)p[5] - p[1]) * i, n);
  } else {
    long long a = (long long)(n - i) * (n - i), b = 2 * (long long)i * (n - i), c = (long long)i * i,
              nn = (long long)n * n;
    x = (int)rdiv64(a * p[0] + b * p[2] + c * p[4], nn);
    y = (int)rdiv64(a * p[1] + b * p[3] + c * p[5], nn);
  }
}

// largest i in [0, n) with off[i] <= j   (off has n + 1 entries, off[0] == 0)
__device__ __forceinline__ uint32_t find_owner(const uint32_t *__restrict__ off, uint32_t n, uint32_t j) {
  uint32_t lo = 0, hi = n;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= j)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

// Fixed-point control points of segment instance j.
__device__ void load_segment(const RenderArgs &a, uint32_t j, int p[6], bool &curve, uint32_t &pid) {
  uint32_t it = find_owner(a.item_seg_off, a.n_items, j);
  uint32_t local = j - __ldg(a.item_seg_off + it);
  const DrawItem &item = a.items[it];
  float m[6];
#pragma unroll
  for (int k = 0; k < 6; k++) m[k] = __ldg(&item.m[k]);
  uint32_t seg_first = __ldg(&item.seg_first);
  double c[6];
  uint32_t pf;
  if (__ldg(&item.is_morph)) {
    const SegMorph &s = a.segs_morph[seg_first + local];
    double r = (double)__ldg(&item.ratio) / 65535.0;
#pragma unroll
    for (int k = 0; k < 6; k++) c[k] = lerp_ref((double)__ldg(&s.s[k]), (double)__ldg(&s.e[k]), r);
    pf = __ldg(&s.path_flags);
  } else {
    const SegStatic &s = a.segs_static[seg_first + local];
#pragma unroll
    for (int k = 0; k < 6; k++) c[k] = (double)__ldg(&s.p[k]);
    pf = __ldg(&s.path_flags);
  }
  curve = (pf >> 31) != 0;
  pid = __ldg(&item.path_off) + (pf & 0x7fffffffu);
  to_device_fx(m, c[0], c[1], p[0], p[1]);
  to_device_fx(m, c[4], c[5], p[4], p[5]);
  if (curve) {
    to_device_fx(m, c[2], c[3], p[2], p[3]);
  } else {
    p[2] = p[0];
    p[3] = p[1];
  }
}

__global__ void k_init(RenderArgs a) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t stride = gridDim.x * blockDim.x;
  if (i == 0) {
    a.totals->n_edges = a.totals->n_slots = a.totals->n_records = 0;
    a.totals->overflow = 0;
    a.totals->error = 0;
    a.totals->work = 0;
  }
  for (uint32_t p = i; p < a.n_paths; p += stride) {
    a.path_bbox[4 * p + 0] = INT_MAX;
    a.path_bbox[4 * p + 1] = INT_MAX;
    a.path_bbox[4 * p + 2] = INT_MIN;
    a.path_bbox[4 * p + 3] = INT_MIN;
  }
}

__global__ void k_flatten_count(RenderArgs a) {
  uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < a.n_seginst; j += stride) {
    int p[6];
    bool curve;
    uint32_t pid;
    load_segment(a, j, p, curve, pid);
    a.seg_edge_off[j] = (uint32_t)piece_count(curve, p);
    int minx = min(p[0], min(p[2], p[4])), maxx = max(p[0], max(p[2], p[4]));
    int miny = min(p[1], min(p[3], p[5])), maxy = max(p[1], max(p[3], p[5]));
    atomicMin(&a.path_bbox[4 * pid + 0], minx);
    atomicMin(&a.path_bbox[4 * pid + 1], miny);
    atomicMax(&a.path_bbox[4 * pid + 2], maxx);
    atomicMax(&a.path_bbox[4 * pid + 3], maxy);
  }
}

__global__ void k_flatten_emit(RenderArgs a) {
  if (a.totals->overflow) return;
  uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < a.n_seginst; j += stride) {
    int p[6];
    bool curve;
    uint32_t pid;
    load_segment(a, j, p, curve, pid);
    uint32_t off = a.seg_edge_off[j];
    int n = (int)(a.seg_edge_off[j + 1] - off);
    int px, py;
    piece_point(curve, p, n, 0, px, py);
    for (int i = 1; i <= n; i++) {
      int qx, qy;
      piece_point(curve, p, n, i, qx, qy);
      a.edges[off + i - 1] = make_int4(px, py, qx, qy);
      a.edge_pid[off + i - 1] = pid;
      px = qx;
      py = qy;
    }
  }
}

// ======================================================================================================
// exclusive scan of a u32 array whose length may only be known on the device
// ======================================================================================================

constexpr int kScanBlocks = kNumSM * 4;
constexpr int kScanThreads = 256;
constexpr int kScanTile = kScanThreads * 4;

__device__ __forceinline__ uint32_t scan_chunk(uint32_t n) {
  uint32_t per = (n + kScanBlocks - 1) / kScanBlocks;
  return ((per + kScanTile - 1) / kScanTile) * kScanTile;
}

__device__ __forceinline__ uint32_t block_reduce(uint32_t v, uint32_t *sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  uint32_t t = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sh[w];
  __syncthreads();
  return t;
}

// n = n_ptr ? min(*n_ptr, n_cap) : n_cap
__global__ void k_scan_partials(const uint32_t *__restrict__ data, const uint32_t *n_ptr, uint32_t n_cap,
                                uint32_t *partials) {
  __shared__ uint32_t sh[32];
  uint32_t n = n_ptr ? min(*n_ptr, n_cap) : n_cap;
  uint32_t chunk = scan_chunk(n);
  uint64_t begin = (uint64_t)blockIdx.x * chunk;
  uint64_t end = min((uint64_t)n, begin + chunk);
  uint32_t s = 0;
  for (uint64_t i = begin + threadIdx.x; i < end; i += blockDim.x) s += data[i];
  s = block_reduce(s, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// Scans the block partials; stores the grand total at data[n] and *total_out; raises overflow_bit when the
// total exceeds total_cap.
__global__ void k_scan_spine(uint32_t *partials, uint32_t *data, const uint32_t *n_ptr, uint32_t n_cap,
                             uint32_t *total_out, uint32_t total_cap, uint32_t *overflow, uint32_t overflow_bit) {
  __shared__ uint32_t sh[kScanBlocks];
  uint32_t n = n_ptr ? min(*n_ptr, n_cap) : n_cap;
  for (int i = threadIdx.x; i < kScanBlocks; i += blockDim.x) sh[i] = partials[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int i = 0; i < kScanBlocks; i++) {
      uint32_t v = sh[i];
      sh[i] = run;
      run += v;
    }
    data[n] = run;
    if (total_out) *total_out = run;
    if (overflow && run > total_cap) atomicOr(overflow, overflow_bit);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kScanBlocks; i += blockDim.x) partials[i] = sh[i];
}

__global__ void k_scan_apply(uint32_t *data, const uint32_t *n_ptr, uint32_t n_cap, const uint32_t *partials) {
  __shared__ uint32_t sh[32];
  __shared__ uint32_t carry_sh;
  uint32_t n = n_ptr ? min(*n_ptr, n_cap) : n_cap;
  uint32_t chunk = scan_chunk(n);
  uint64_t begin = (uint64_t)blockIdx.x * chunk;
  uint64_t end = min((uint64_t)n, begin + chunk);
  if (threadIdx.x == 0) carry_sh = partials[blockIdx.x];
  __syncthreads();
  for (uint64_t base = begin; base < end; base += kScanTile) {
    uint64_t i0 = base + (uint64_t)threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (i0 + k < end) ? data[i0 + k] : 0u;
    uint32_t tsum = v[0] + v[1] + v[2] + v[3];
    // inclusive warp scan of thread sums
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) wbase += sh[w];
    uint32_t tile_total = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) tile_total += sh[w];
    uint32_t ex = carry_sh + wbase + inc - tsum;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (i0 + k < end) data[i0 + k] = ex;
      ex += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_sh += tile_total;
    __syncthreads();
  }
}

// ======================================================================================================
// path setup: tile bbox, slot allocation size, paint instance   
// ======================================================================================================

__device__ uint32_t solid_premul8(double r8, double g8, double b8, double alpha) {
  double af = (double)(float)alpha;
  if (af < 0) af = 0;
  if (af > 1) af = 1;
  uint32_t a16 = (uint32_t)(af * 65535.0 + 0.5);
  uint32_t r16 = (uint32_t)(((r8 / 255.0) * af) * 65535.0 + 0.5);
  uint32_t g16 = (uint32_t)(((g8 / 255.0) * af) * 65535.0 + 0.5);
  uint32_t b16 = (uint32_t)(((b8 / 255.0) * af) * 65535.0 + 0.5);
  return (r16 >> 8) | ((g16 >> 8) << 8) | ((b16 >> 8) << 16) | ((a16 >> 8) << 24);
}

__device__ uint32_t morph_solid(const uint8_t *c0, const uint8_t *c1, double r) {
  double ch[4];
  for (int i = 0; i < 4; i++) ch[i] = lerp_ref(c0[i] / 255.0, c1[i] / 255.0, r);
  double red = (double)(((long long)(ch[0] * 255.0)) & 0xff);
  double g = ceil((double)(float)(ch[1] * 255.0));
  double b = ceil((double)(float)(ch[2] * 255.0));
  if (g < 0) g = 0;
  if (g > 255) g = 255;
  if (b < 0) b = 0;
  if (b > 255) b = 255;
  return solid_premul8(red, g, b, ch[3]);
}

__global__ void k_path_setup(RenderArgs a) {
  uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x; pid < a.n_paths; pid += stride) {
    uint32_t it = find_owner(a.item_path_off, a.n_items, pid);
    const DrawItem &item = a.items[it];
    const DefPaint &dp = a.def_paints[item.paint_first + (pid - a.item_path_off[it])];
    // ---- tile bbox ----
    int minx = a.path_bbox[4 * pid + 0], miny = a.path_bbox[4 * pid + 1];
    int maxx = a.path_bbox[4 * pid + 2], maxy = a.path_bbox[4 * pid + 3];
    int bx0 = 0, by0 = 0, bw = 0, bh = 0;
    if (minx <= maxx) {
      bx0 = max(minx >> 12, 0);
      by0 = max(miny >> 12, 0);
      int bx1 = min(maxx >> 12, a.tiles_x - 1), by1 = min(maxy >> 12, a.tiles_y - 1);
      if (bx1 >= bx0 && by1 >= by0) {
        bw = bx1 - bx0 + 1;
        bh = by1 - by0 + 1;
      }
    }
    // ---- paint ----
    PathRec rec;
    rec.color = 0;
    uint32_t flags = 0;
    bool valid = true;
    double ratio = (double)item.ratio / 65535.0;
    if (dp.type == PAINT_SOLID) {
      if (dp.flags & PF_COLOR_MORPH)
        rec.color = morph_solid(dp.color0, dp.color1, ratio);
      else
        rec.color = solid_premul8(dp.color0[0], dp.color0[1], dp.color0[2], dp.color0[3] / 255.0);
      if ((rec.color >> 24) == 255) flags |= 1u;
      if (rec.color == 0) valid = false;  // composites nothing
    } else {
      double ma = item.m[0], md = item.m[1], mb = item.m[2], mc = item.m[3], mtx = item.m[4], mty = item.m[5];
      double fa = dp.matrix[0], fb = dp.matrix[1], fc = dp.matrix[2], fd = dp.matrix[3], ftx = dp.matrix[4],
             fty = dp.matrix[5];
      double ca = (ma * fa + mc * fb) / 20.0;
      double cb = (mb * fa + md * fb) / 20.0;
      double cc = (ma * fc + mc * fd) / 20.0;
      double cd = (mb * fc + md * fd) / 20.0;
      double ctx = ((ma * ftx + mc * fty) + mtx) / 20.0;
      double cty = ((mb * ftx + md * fty) + mty) / 20.0;
      double det = ca * cd - cb * cc;
      PaintInst pi;
      memset(&pi, 0, sizeof pi);
      if (!(det != 0.0) || isinf(det) || isnan(det)) {
        valid = false;
      } else {
        double ia = cd / det, ib = -cb / det, ic = -cc / det, id = ca / det;
        double itx = -(ia * ctx + ic * cty), ity = -(ib * ctx + id * cty);
        pi.inv[0] = (float)ia;
        pi.inv[1] = (float)ib;
        pi.inv[2] = (float)ic;
        pi.inv[3] = (float)id;
        pi.inv[4] = (float)itx;
        pi.inv[5] = (float)ity;
        pi.spread = dp.spread;
        pi.repeating = dp.repeating;
        if (dp.type == PAINT_BITMAP) {
          const BitmapDev &bm = a.bitmaps[dp.bitmap_id & 0xffff];
          if (!bm.valid) {
            atomicOr(&a.totals->error, 1u);  // BitmapNotFound (node-canvas-bitmap-service.ts:41-43)
            valid = false;
          } else {
            pi.ptr = bm.tex;
            pi.bw = bm.w;
            pi.bh = bm.h;
            double dx = sqrt(ia * ia + ic * ic), dy = sqrt(ib * ib + id * id);
            if (dx > 16.0) dx = 16.0;
            if (dy > 16.0) dy = 16.0;
            if (dx < 1.0 / 0.75) dx = 1.0;
            if (dy < 1.0 / 0.75) dy = 1.0;
            pi.rx = (float)dx;
            pi.ry = (float)dy;
            if (bm.opaque && dp.repeating) flags |= 1u;
          }
        } else {
          double fp = dp.type == PAINT_FOCAL ? dp.focal : 0.0;
          if (fp > 0.98) fp = 0.98;
          if (fp < -0.98) fp = -0.98;
          pi.focal = (float)fp;
          pi.omf = (float)(1.0 - fp * fp);
          pi.ptr = (unsigned long long)(a.ramps + (size_t)dp.lut * 257 * 4);
          if (dp.flags & PF_OPAQUE_RAMP) flags |= 1u;
        }
      }
      a.paint_inst[pid] = pi;
    }
    if (!valid) bw = bh = 0;
    rec.xy0 = (uint32_t)bx0 | ((uint32_t)by0 << 16);
    rec.wh = (uint32_t)bw | ((uint32_t)bh << 16);
    rec.info = dp.type | (flags << 8);
    a.path_rec[pid] = rec;
    a.path_slot_off[pid] = (uint32_t)(bw * bh);
  }
}

__global__ void k_zero_slots(RenderArgs a) {
  if (a.totals->overflow) return;
  uint32_t n = a.totals->n_slots;
  uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    a.slot_count[i] = 0;
    a.slot_backdrop[i] = 0;
  }
}

__global__ void k_copy_counts(RenderArgs a) {
  if (a.totals->overflow) return;
  uint32_t n = a.totals->n_slots;
  uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a.slot_off[i] = a.slot_count[i];
}

__global__ void k_zero_cursor(RenderArgs a) {
  if (a.totals->overflow) return;
  uint32_t n = a.totals->n_slots;
  uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a.slot_count[i] = 0;
}

// ======================================================================================================
// K2: tile binning (count, then scatter of tile-clipped 8-byte records)
// ======================================================================================================

__device__ __forceinline__ int xat(int x0, int y0, int x1, int y1, int Y) {
  return x0 + (int)rdiv64(((long long)Y - y0) * ((long long)x1 - x0), (long long)y1 - y0);
}

__device__ __forceinline__ unsigned long long pack_record(int xa, int ya, int xb, int yb, int fs, int fe) {
  return (unsigned long long)xa | ((unsigned long long)ya << 13) | ((unsigned long long)xb << 26) |
         ((unsigned long long)yb << 39) | ((unsigned long long)fs << 52) | ((unsigned long long)fe << 53);
}

template <int MODE>  // 0 = count + backdrop deltas, 1 = scatter records
__device__ void bin_edge(const RenderArgs &a, int x0, int y0, int x1, int y1, int bx0, int by0, int bw, int bh,
                         uint32_t slot_base) {
  const int B = kTileFx;
  int b_first, b_last;
  if (y0 == y1) {
    b_first = b_last = y0 >> 12;
  } else {
    int ylo = min(y0, y1), yhi = max(y0, y1);
    b_first = ylo >> 12;
    b_last = (yhi - 1) >> 12;
  }
  // rows outside the path's grid are outside the viewport (the bbox covers every edge of the path)
  b_first = max(b_first, by0);
  b_last = min(b_last, by0 + bh - 1);
  for (int b = b_first; b <= b_last; b++) {
    int Yt = b * B, Yb = Yt + B;
    int xs, ys, xe, ye;
    if (y0 == y1) {
      xs = x0, ys = y0, xe = x1, ye = y1;
    } else if (y0 < y1) {
      ys = max(y0, Yt);
      ye = min(y1, Yb);
      xs = (ys == y0) ? x0 : xat(x0, y0, x1, y1, ys);
      xe = (ye == y1) ? x1 : xat(x0, y0, x1, y1, ye);
    } else {
      ys = min(y0, Yb);
      ye = max(y1, Yt);
      xs = (ys == y0) ? x0 : xat(x0, y0, x1, y1, ys);
      xe = (ye == y1) ? x1 : xat(x0, y0, x1, y1, ye);
    }
    uint32_t row_base = slot_base + (uint32_t)((b - by0) * bw);
    if (MODE == 0) {
      if (ys == Yt) {
        int lx = max((xs >> 12) + 1 - bx0, 0);
        if (lx < bw) atomicAdd(&a.slot_backdrop[row_base + lx], 1);
      }
      if (ye == Yt) {
        int lx = max((xe >> 12) + 1 - bx0, 0);
        if (lx < bw) atomicAdd(&a.slot_backdrop[row_base + lx], -1);
      }
    }
    int c0 = max(min(xs, xe) >> 12, bx0), c1 = min(max(xs, xe) >> 12, bx0 + bw - 1);
    for (int t = c0; t <= c1; t++) {
      int X0 = t * B, X1 = X0 + B;
      int ax, ay, bx, by, fs = 0, fe = 0;
      if (xs < X0) {
        ax = X0, ay = ys + (int)rdiv64(((long long)X0 - xs) * ((long long)ye - ys), (long long)xe - xs), fs = 1;
      } else if (xs > X1) {
        ax = X1, ay = ys + (int)rdiv64(((long long)X1 - xs) * ((long long)ye - ys), (long long)xe - xs);
      } else {
        ax = xs, ay = ys;
      }
      if (xe < X0) {
        bx = X0, by = ys + (int)rdiv64(((long long)X0 - xs) * ((long long)ye - ys), (long long)xe - xs), fe = 1;
      } else if (xe > X1) {
        bx = X1, by = ys + (int)rdiv64(((long long)X1 - xs) * ((long long)ye - ys), (long long)xe - xs);
      } else {
        bx = xe, by = ye;
      }
      if (ay == by && !fs && !fe) continue;
      uint32_t slot = row_base + (uint32_t)(t - bx0);
      if (MODE == 0) {
        atomicAdd(&a.slot_count[slot], 1u);
      } else {
        uint32_t pos = a.slot_off[slot] + atomicAdd(&a.slot_count[slot], 1u);
        a.records[pos] = pack_record(ax - X0, ay - Yt, bx - X0, by - Yt, fs, fe);
      }
    }
  }
}

template <int MODE>
__global__ void k_bin(RenderArgs a) {
  if (a.totals->overflow) return;
  uint32_t n = a.totals->n_edges;
  uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    int4 ed = a.edges[e];
    uint32_t pid = a.edge_pid[e];
    PathRec rec = a.path_rec[pid];
    int bw = rec.wh & 0xffff, bh = rec.wh >> 16;
    if (bw == 0) continue;
    bin_edge<MODE>(a, ed.x, ed.y, ed.z, ed.w, rec.xy0 & 0xffff, rec.xy0 >> 16, bw, bh, a.path_slot_off[pid]);
  }
}

// prefix sum of the backdrop deltas along each tile row of each path grid: one warp per path, lanes over rows
__global__ void k_backdrop(RenderArgs a) {
  if (a.totals->overflow) return;
  uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t pid = warp; pid < a.n_paths; pid += nwarps) {
    uint32_t wh = a.path_rec[pid].wh;
    int bw = wh & 0xffff, bh = wh >> 16;
    if (bw <= 1) continue;
    uint32_t base = a.path_slot_off[pid];
    for (int row = lane; row < bh; row += 32) {
      int32_t *p = a.slot_backdrop + base + (uint32_t)row * bw;
      int run = 0;
      for (int x = 0; x < bw; x++) {
        run += p[x];
        p[x] = run;
      }
    }
  }
}

// ======================================================================================================
// K3 + K4: per-tile coverage, paint evaluation and pixman-style OVER.  One warp per 16x16 tile.
// ======================================================================================================

constexpr int kFineWarps = 8;
constexpr int kAccStride = 20;  // words per accumulator row: conflict-free 128-bit row reads

__device__ __forceinline__ uint32_t mul_un8x4(uint32_t p, uint32_t m) {
  uint32_t rb = (p & 0x00ff00ffu) * m + 0x00800080u;
  rb = ((rb + ((rb >> 8) & 0x00ff00ffu)) >> 8) & 0x00ff00ffu;
  uint32_t ag = ((p >> 8) & 0x00ff00ffu) * m + 0x00800080u;
  ag = (ag + ((ag >> 8) & 0x00ff00ffu)) & 0xff00ff00u;
  return rb | ag;
}

__device__ __forceinline__ uint32_t over_masked(uint32_t dst, uint32_t src, uint32_t m) {
  uint32_t s = (m == 255u) ? src : mul_un8x4(src, m);
  uint32_t sa = s >> 24;
  if (sa == 255u) return s;
  if (s == 0u) return dst;
  return s + mul_un8x4(dst, 255u - sa);
}

// Adds one record's signed-area contribution for pixel row `r` into acc_row[0..15] as prefix differences.
__device__ __forceinline__ void accumulate_row(int xa, int ya, int xb, int yb, int r, int *acc_row) {
  int ylo = min(ya, yb), yhi = max(ya, yb);
  int yt = max(ylo, 256 * r), ybm = min(yhi, 256 * (r + 1));
  if (ybm <= yt) return;
  const float k = 1.0f / 256.0f;
  int s = yb > ya ? 1 : -1;
  int D = s * (ybm - yt) * 256;
  float Df = (float)D;
  float xaf = (float)xa * k, yaf = (float)ya * k, xbf = (float)xb * k, ybf = (float)yb * k;
  float slope = (xbf - xaf) / (ybf - yaf);
  float xlo = fminf(xaf, xbf), xhi = fmaxf(xaf, xbf);
  float ytf = (float)yt * k, ybmf = (float)ybm * k;
  float t0 = (ytf - yaf) * slope;
  float t1 = (ybmf - yaf) * slope;
  float xt = xaf + t0, xm = xaf + t1;
  xt = fminf(fmaxf(xt, xlo), xhi);
  xm = fminf(fmaxf(xm, xlo), xhi);
  float xmin = fminf(xt, xm), xmax = fmaxf(xt, xm);
  float w = xmax - xmin;
  float inv2w = w > 0.0f ? 0.5f / w : 0.0f;
  int i0 = (int)floorf(xmin);   // first pixel that is not entirely left of the edge
  int iend = (int)ceilf(xmax);  // first pixel entirely right of the edge
  int prev = 0;
  for (int i = i0; i < iend && i < 16; i++) {
    float fi = (float)i, fi1 = (float)(i + 1);
    float u0 = fmaxf(fi - xmin, 0.0f);
    float u1 = fminf(fi1 - xmin, w);
    float a0 = (u0 * u0) * inv2w;
    float a1 = (u1 * u1) * inv2w + fmaxf(fi1 - xmax, 0.0f);
    float f = a1 - a0;
    f = fminf(fmaxf(f, 0.0f), 1.0f);
    int c = __float2int_rn(Df * f);
    atomicAdd(&acc_row[i], c - prev);
    prev = c;
  }
  if (iend < 16) atomicAdd(&acc_row[iend], D - prev);
}

__device__ __forceinline__ int floormod(int a, int n) {
  int r = a % n;
  return r < 0 ? r + n : r;
}

__device__ uint32_t eval_paint(uint32_t type, const PaintInst &p, int X, int Y) {
  float xc = (float)X + 0.5f, yc = (float)Y + 0.5f;
  float gx = (p.inv[0] * xc + p.inv[2] * yc) + p.inv[4];
  float gy = (p.inv[1] * xc + p.inv[3] * yc) + p.inv[5];
  if (type == PAINT_BITMAP) {
    cudaTextureObject_t tex = (cudaTextureObject_t)p.ptr;
    float hrx = p.rx * 0.5f, hry = p.ry * 0.5f;
    float irx = 1.0f / p.rx, iry = 1.0f / p.ry;
    float lox = gx - hrx, hix = gx + hrx, loy = gy - hry, hiy = gy + hry;
    int i0 = (int)floorf(lox), j0 = (int)floorf(loy);
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    for (int j = j0; (float)j < hiy; j++) {
      float wl = fmaxf(loy, (float)j), wh = fminf(hiy, (float)(j + 1));
      float wy = fmaxf(wh - wl, 0.0f) * iry;
      int jj = j;
      if (p.repeating)
        jj = floormod(j, p.bh);
      else if (j < 0 || j >= p.bh)
        continue;
      for (int i = i0; (float)i < hix; i++) {
        float vl = fmaxf(lox, (float)i), vh = fminf(hix, (float)(i + 1));
        float wx = fmaxf(vh - vl, 0.0f) * irx;
        int ii = i;
        if (p.repeating)
          ii = floormod(i, p.bw);
        else if (i < 0 || i >= p.bw)
          continue;
        uchar4 t = tex2D<uchar4>(tex, (float)ii + 0.5f, (float)jj + 0.5f);
        float wgt = wx * wy;
        acc0 = acc0 + wgt * (float)t.x;
        acc1 = acc1 + wgt * (float)t.y;
        acc2 = acc2 + wgt * (float)t.z;
        acc3 = acc3 + wgt * (float)t.w;
      }
    }
    uint32_t o = (uint32_t)__float2int_rn(fminf(fmaxf(acc0, 0.0f), 255.0f));
    o |= (uint32_t)__float2int_rn(fminf(fmaxf(acc1, 0.0f), 255.0f)) << 8;
    o |= (uint32_t)__float2int_rn(fminf(fmaxf(acc2, 0.0f), 255.0f)) << 16;
    o |= (uint32_t)__float2int_rn(fminf(fmaxf(acc3, 0.0f), 255.0f)) << 24;
    return o;
  }
  float t;
  if (type == PAINT_LINEAR) {
    t = gx * (1.0f / 32768.0f) + 0.5f;
  } else {
    float nx = gx * (1.0f / 16384.0f), ny = gy * (1.0f / 16384.0f);
    float dx = nx - p.focal;
    float a = dx * dx;
    float b = ny * ny;
    float c = p.omf * b;
    float disc = a + c;
    float s = sqrtf(disc);
    float num = p.focal * dx + s;
    t = num / p.omf;
  }
  if (p.spread == SWFR_SPREAD_PAD) {
    t = fminf(fmaxf(t, 0.0f), 1.0f);
  } else if (p.spread == SWFR_SPREAD_REPEAT) {
    t = t - floorf(t);
    t = fminf(fmaxf(t, 0.0f), 1.0f);
  } else {
    float u = t * 0.5f;
    u = u - floorf(u);
    u = u * 2.0f;
    t = u > 1.0f ? 2.0f - u : u;
    t = fminf(fmaxf(t, 0.0f), 1.0f);
  }
  float pos = t * 256.0f;
  int i = (int)floorf(pos);
  i = min(max(i, 0), 255);
  float fr = pos - (float)i;
  const float4 *lut = (const float4 *)p.ptr;
  float4 l0 = __ldg(lut + i), l1 = __ldg(lut + i + 1);
  float v0 = l0.x + (l1.x - l0.x) * fr;
  float v1 = l0.y + (l1.y - l0.y) * fr;
  float v2 = l0.z + (l1.z - l0.z) * fr;
  float A = l0.w + (l1.w - l0.w) * fr;
  uint32_t o = (uint32_t)__float2int_rn(fminf(fmaxf(A * 255.0f, 0.0f), 255.0f)) << 24;
  o |= (uint32_t)__float2int_rn(fminf(fmaxf((v0 * A) * 255.0f, 0.0f), 255.0f));
  o |= (uint32_t)__float2int_rn(fminf(fmaxf((v1 * A) * 255.0f, 0.0f), 255.0f)) << 8;
  o |= (uint32_t)__float2int_rn(fminf(fmaxf((v2 * A) * 255.0f, 0.0f), 255.0f)) << 16;
  return o;
}

struct Probe {
  uint32_t o0, o1;
  int bd;
  uint32_t info, color;
  bool hit;
};

__device__ __forceinline__ Probe probe_slot(const RenderArgs &a, uint32_t pid, bool valid, int tx, int ty) {
  Probe pr;
  pr.hit = false;
  pr.o0 = pr.o1 = 0;
  pr.bd = 0;
  pr.info = pr.color = 0;
  if (!valid) return pr;
  uint4 rc = __ldg(reinterpret_cast<const uint4 *>(a.path_rec + pid));
  int bx0 = rc.x & 0xffff, by0 = rc.x >> 16, bw = rc.y & 0xffff, bh = rc.y >> 16;
  int lx = tx - bx0, ly = ty - by0;
  if (lx < 0 || ly < 0 || lx >= bw || ly >= bh) return pr;
  uint32_t slot = __ldg(a.path_slot_off + pid) + (uint32_t)(ly * bw + lx);
  pr.o0 = __ldg(a.slot_off + slot);
  pr.o1 = __ldg(a.slot_off + slot + 1);
  pr.bd = __ldg(a.slot_backdrop + slot);
  pr.info = rc.z;
  pr.color = rc.w;
  pr.hit = (pr.o1 > pr.o0) || (pr.bd != 0);
  return pr;
}

__global__ void __launch_bounds__(kFineWarps * 32) k_fine(RenderArgs a) {
  if (a.totals->overflow) return;
  __shared__ int acc_sh[kFineWarps][16 * kAccStride];
  __shared__ int cross_sh[kFineWarps][20];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int *acc = acc_sh[warp];
  int *cross = cross_sh[warp];
  for (int i = lane; i < 16 * kAccStride; i += 32) acc[i] = 0;
  if (lane < 20) cross[lane] = 0;
  __syncwarp();
  const int row = lane & 15, half = lane >> 4;
  const uint32_t tiles = (uint32_t)(a.tiles_x * a.tiles_y);
  const uint32_t total = tiles * a.n_frames;
  while (true) {
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(&a.totals->work, 1u);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= total) break;
    uint32_t frame = w / tiles, tile = w - frame * tiles;
    int ty = (int)(tile / (uint32_t)a.tiles_x), tx = (int)(tile - (uint32_t)ty * a.tiles_x);
    uint32_t p_begin = __ldg(a.frame_path_off + frame), p_end = __ldg(a.frame_path_off + frame + 1);

    // ---- pass 1: the last opaque full-tile cover hides everything painted before it ----
    uint32_t start = p_begin;
    for (uint32_t hi = p_end; hi > p_begin;) {
      uint32_t lo = hi - p_begin >= 32 ? hi - 32 : p_begin;
      uint32_t pid = lo + lane;
      Probe pr = probe_slot(a, pid, pid < hi, tx, ty);
      bool cover = pr.hit && pr.o1 == pr.o0 && ((pr.info >> 8) & 1u);
      uint32_t mask = __ballot_sync(0xffffffffu, cover);
      if (mask) {
        start = lo + (31 - __clz(mask));
        break;
      }
      hi = lo;
    }

    // ---- pass 2: composite in paint order ----
    uint32_t px[8];
#pragma unroll
    for (int i = 0; i < 8; i++) px[i] = 0;
    const int X0 = tx * kTile + half * 8, Y = ty * kTile + row;
    for (uint32_t base = start; base < p_end; base += 32) {
      uint32_t pid = base + lane;
      Probe pr = probe_slot(a, pid, pid < p_end, tx, ty);
      uint32_t mask = __ballot_sync(0xffffffffu, pr.hit);
      while (mask) {
        int src_lane = __ffs(mask) - 1;
        mask &= mask - 1;
        uint32_t o0 = __shfl_sync(0xffffffffu, pr.o0, src_lane);
        uint32_t o1 = __shfl_sync(0xffffffffu, pr.o1, src_lane);
        int bd = __shfl_sync(0xffffffffu, pr.bd, src_lane);
        uint32_t info = __shfl_sync(0xffffffffu, pr.info, src_lane);
        uint32_t color = __shfl_sync(0xffffffffu, pr.color, src_lane);
        uint32_t cur_pid = base + src_lane;
        uint32_t type = info & 0xff;
        uint32_t m[8];
        if (o1 == o0) {
#pragma unroll
          for (int i = 0; i < 8; i++) m[i] = 255u;
        } else {
          uint32_t nrec = o1 - o0;
          bool any_cross = false;
          if (nrec >= 16) {
            // many records: one lane per record, rows in a loop
            for (uint32_t k = lane; k < nrec; k += 32) {
              unsigned long long rc = __ldg(a.records + o0 + k);
              int xa = (int)(rc & 0x1fff), ya = (int)((rc >> 13) & 0x1fff), xb = (int)((rc >> 26) & 0x1fff),
                  yb = (int)((rc >> 39) & 0x1fff);
              int fs = (int)((rc >> 52) & 1), fe = (int)((rc >> 53) & 1);
              if (fs | fe) {
                int yc = fs ? ya : yb, sgn = fs ? -1 : 1;
                int r0 = min(yc >> 8, 15);
                int hq = min(max((r0 + 1) * 256 - yc, 0), 256) * 256;
                atomicAdd(&cross[r0], sgn * hq);
                atomicAdd(&cross[r0 + 1], sgn * (65536 - hq));
                any_cross = true;
              }
              if (ya != yb) {
                int ylo = min(ya, yb), yhi = max(ya, yb);
                for (int r = ylo >> 8; r <= ((yhi - 1) >> 8); r++) accumulate_row(xa, ya, xb, yb, r, acc + r * kAccStride);
              }
            }
          } else {
            // few records: one lane per (record, row)
            for (uint32_t k = lane; k < nrec * 16; k += 32) {
              unsigned long long rc = __ldg(a.records + o0 + (k >> 4));
              int r = (int)(k & 15);
              int xa = (int)(rc & 0x1fff), ya = (int)((rc >> 13) & 0x1fff), xb = (int)((rc >> 26) & 0x1fff),
                  yb = (int)((rc >> 39) & 0x1fff);
              int fs = (int)((rc >> 52) & 1), fe = (int)((rc >> 53) & 1);
              if (fs | fe) {
                int yc = fs ? ya : yb, sgn = fs ? -1 : 1;
                int r0 = min(yc >> 8, 15);
                if (r == r0) {  // one lane per record posts the crossing term
                  int hq = min(max((r0 + 1) * 256 - yc, 0), 256) * 256;
                  atomicAdd(&cross[r0], sgn * hq);
                  atomicAdd(&cross[r0 + 1], sgn * (65536 - hq));
                }
                any_cross = true;
              }
              if (ya != yb) accumulate_row(xa, ya, xb, yb, r, acc + r * kAccStride);
            }
          }
          any_cross = __any_sync(0xffffffffu, any_cross);
          __syncwarp();
          // row prefix: 8 accumulators of this lane, carry from the left half
          int v[8];
          int4 q0 = *reinterpret_cast<int4 *>(acc + row * kAccStride + half * 8);
          int4 q1 = *reinterpret_cast<int4 *>(acc + row * kAccStride + half * 8 + 4);
          v[0] = q0.x, v[1] = q0.y, v[2] = q0.z, v[3] = q0.w, v[4] = q1.x, v[5] = q1.y, v[6] = q1.z, v[7] = q1.w;
#pragma unroll
          for (int i = 1; i < 8; i++) v[i] += v[i - 1];
          int left = __shfl_sync(0xffffffffu, v[7], lane & 15);
          int basev = bd * 65536 + (half ? left : 0);
          if (any_cross) {
            int cr = cross[lane & 15];
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
              int t = __shfl_up_sync(0xffffffffu, cr, o, 16);
              if ((lane & 15) >= o) cr += t;
            }
            basev += cr;
          }
          __syncwarp();
          *reinterpret_cast<int4 *>(acc + row * kAccStride + half * 8) = make_int4(0, 0, 0, 0);
          *reinterpret_cast<int4 *>(acc + row * kAccStride + half * 8 + 4) = make_int4(0, 0, 0, 0);
          if (any_cross && lane < 20) cross[lane] = 0;
#pragma unroll
          for (int i = 0; i < 8; i++) {
            int s = v[i] + basev;
            s = s < 0 ? -s : s;
            s = min(s, 65536);
            m[i] = ((uint32_t)s * 255u + 32768u) >> 16;
          }
          __syncwarp();
        }
        // ---- paint + blend ----
        if (type == PAINT_SOLID) {
#pragma unroll
          for (int i = 0; i < 8; i++)
            if (m[i]) px[i] = over_masked(px[i], color, m[i]);
        } else {
          const PaintInst &pi = a.paint_inst[cur_pid];
#pragma unroll 1
          for (int i = 0; i < 8; i++)
            if (m[i]) px[i] = over_masked(px[i], eval_paint(type, pi, X0 + i, Y), m[i]);
        }
      }
    }

    // ---- store: 8 pixels = 32 bytes per lane ----
    if (Y < a.height) {
      uint32_t *dst = a.frames + ((size_t)frame * a.height + Y) * a.width + X0;
      if (X0 + 8 <= a.width && (a.width & 3) == 0) {
        reinterpret_cast<uint4 *>(dst)[0] = make_uint4(px[0], px[1], px[2], px[3]);
        reinterpret_cast<uint4 *>(dst)[1] = make_uint4(px[4], px[5], px[6], px[7]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++)
          if (X0 + i < a.width) dst[i] = px[i];
      }
    }
  }
}

// ======================================================================================================
// readback helpers and debug taps
// ======================================================================================================

__global__ void k_unpremultiply(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, uint64_t n) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint32_t p = src[i], al = p >> 24, o = 0;
    if (al) {
      uint32_t r = ((p & 255) * 255u + al / 2) / al, g = (((p >> 8) & 255) * 255u + al / 2) / al,
               b = (((p >> 16) & 255) * 255u + al / 2) / al;
      o = min(r, 255u) | (min(g, 255u) << 8) | (min(b, 255u) << 16) | (al << 24);
    }
    dst[i] = o;
  }
}

__global__ void k_premultiply(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, uint64_t n) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint32_t p = src[i], al = p >> 24;
    uint32_t r = ((p & 255) * al + 127u) / 255u, g = (((p >> 8) & 255) * al + 127u) / 255u,
             b = (((p >> 16) & 255) * al + 127u) / 255u;
    dst[i] = r | (g << 8) | (b << 16) | (al << 24);
  }
}

__global__ void k_tile_counts(RenderArgs a, uint32_t frame, uint32_t *counts) {
  uint32_t p_begin = a.frame_path_off[frame], p_end = a.frame_path_off[frame + 1];
  uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t pid = p_begin + blockIdx.x * blockDim.x + threadIdx.x; pid < p_end; pid += stride) {
    PathRec rec = a.path_rec[pid];
    int bx0 = rec.xy0 & 0xffff, by0 = rec.xy0 >> 16, bw = rec.wh & 0xffff, bh = rec.wh >> 16;
    uint32_t base = a.path_slot_off[pid];
    for (int y = 0; y < bh; y++)
      for (int x = 0; x < bw; x++) {
        uint32_t s = base + (uint32_t)(y * bw + x);
        uint32_t c = a.slot_off[s + 1] - a.slot_off[s];
        if (c) atomicAdd(&counts[(by0 + y) * a.tiles_x + bx0 + x], c);
      }
  }
}

}  // namespace

// ======================================================================================================
// launchers
// ======================================================================================================

static void scan_u32(uint32_t *data, const uint32_t *n_ptr, uint32_t n_cap, uint32_t *tmp, uint32_t *total_out,
                     uint32_t total_cap, uint32_t *overflow, uint32_t bit, cudaStream_t st, int &launches) {
  k_scan_partials<<<kScanBlocks, kScanThreads, 0, st>>>(data, n_ptr, n_cap, tmp);
  k_scan_spine<<<1, 256, 0, st>>>(tmp, data, n_ptr, n_cap, total_out, total_cap, overflow, bit);
  k_scan_apply<<<kScanBlocks, kScanThreads, 0, st>>>(data, n_ptr, n_cap, tmp);
  launches += 3;
}

const char *stage_name(int i) {
  static const char *names[kNumStages] = {"flatten_count", "scan_edges",   "path_setup",  "flatten_emit",
                                          "bin_count",     "scan_records", "bin_scatter", "fine"};
  return (i >= 0 && i < kNumStages) ? names[i] : "?";
}

// ev (optional): kNumStages + 1 events recorded at the stage boundaries (profiling runs only).
int launch_render(const RenderArgs &a, cudaStream_t st, cudaEvent_t *ev) {
  int launches = 0;
  const int T = 256;
  auto grid_for = [&](uint64_t n) {
    uint64_t b = (n + T - 1) / T;
    uint64_t cap = (uint64_t)kNumSM * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
  };
  auto mark = [&](int i) {
    if (ev) cudaEventRecord(ev[i], st);
  };
  const unsigned wide = kNumSM * 16;
  mark(0);
  k_init<<<grid_for(a.n_paths), T, 0, st>>>(a);
  launches++;
  if (a.n_seginst) {
    k_flatten_count<<<grid_for(a.n_seginst), T, 0, st>>>(a);
    launches++;
  }
  mark(1);
  // edges: seg_edge_off (piece counts) -> exclusive offsets, total -> totals.n_edges
  scan_u32(a.seg_edge_off, nullptr, a.n_seginst, a.scan_tmp, &a.totals->n_edges, a.caps.edges, &a.totals->overflow, 1u,
           st, launches);
  mark(2);
  if (a.n_paths) {
    k_path_setup<<<grid_for(a.n_paths), T, 0, st>>>(a);
    launches++;
  }
  scan_u32(a.path_slot_off, nullptr, a.n_paths, a.scan_tmp, &a.totals->n_slots, a.caps.slots, &a.totals->overflow, 2u,
           st, launches);
  k_zero_slots<<<wide, T, 0, st>>>(a);
  launches++;
  mark(3);
  if (a.n_seginst) {
    k_flatten_emit<<<grid_for(a.n_seginst), T, 0, st>>>(a);
    launches++;
  }
  mark(4);
  if (a.n_seginst) {
    k_bin<0><<<wide, T, 0, st>>>(a);
    k_backdrop<<<grid_for((uint64_t)a.n_paths * 32), T, 0, st>>>(a);
    launches += 2;
  }
  mark(5);
  // records: slot_count -> slot_off (separate array so counts can be reused as scatter cursors)
  k_copy_counts<<<wide, T, 0, st>>>(a);
  launches++;
  scan_u32(a.slot_off, &a.totals->n_slots, a.caps.slots, a.scan_tmp, &a.totals->n_records, a.caps.records,
           &a.totals->overflow, 4u, st, launches);
  if (a.n_seginst) {
    k_zero_cursor<<<wide, T, 0, st>>>(a);
    launches++;
  }
  mark(6);
  if (a.n_seginst) {
    k_bin<1><<<wide, T, 0, st>>>(a);
    launches++;
  }
  mark(7);
  k_fine<<<kNumSM * 4, kFineWarps * 32, 0, st>>>(a);
  launches++;
  mark(8);
  return launches;
}

void launch_unpremultiply(const uint32_t *src, uint32_t *dst, uint64_t n_px, cudaStream_t st) {
  k_unpremultiply<<<kNumSM * 8, 256, 0, st>>>(src, dst, n_px);
}
void launch_premultiply(const uint32_t *src, uint32_t *dst, uint64_t n_px, cudaStream_t st) {
  k_premultiply<<<kNumSM * 8, 256, 0, st>>>(src, dst, n_px);
}
void launch_tile_counts(const RenderArgs &a, uint32_t frame, uint32_t *counts, cudaStream_t st) {
  k_tile_counts<<<kNumSM * 2, 256, 0, st>>>(a, frame, counts);
}

}  // namespace swfr
