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
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "kernels.h"

namespace swfr {
namespace {

// ======================================================================================================
// programmatic dependent launch (sm_90+): every kernel of a render is launched with the programmatic stream
// serialization attribute, so that its blocks may be scheduled while the tail of its predecessor is still running.
// pdl_enter() is the first statement of such a kernel: it lets the successor be scheduled early in turn
// (launch_dependents) and then waits until the predecessor grid has completed and its writes are visible (wait) -
// nothing before it may touch global memory.  Launched without the attribute both instructions do nothing.
// ======================================================================================================
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;");
  pdl_wait();
}

// ======================================================================================================
// integer helpers (same definitions as the oracle)
// ======================================================================================================

__device__ __forceinline__ long long floordiv64(long long a, long long b) {  // b > 0
  long long q = a / b;
  if ((a % b != 0) && (a < 0)) q -= 1;
  return q;
}
// round half up of a / b (b != 0): the exact integer floor((2a + b) / (2b)), computed without the 64-bit
// integer divide subroutine.  Operands that fit 30/24 bits (every edge of ordinary size) take a float
// reciprocal estimate; larger ones take an FP64 quotient; both are then corrected against the exact integer
// remainder, so the result is the mathematical floor in all cases (same value as the oracle's rdiv64).
__device__ __forceinline__ long long rdiv64(long long a, long long b) {
  if (b < 0) {
    a = -a;
    b = -b;
  }
  long long num = 2 * a + b, den = 2 * b;
  long long an = num < 0 ? -num : num;
  if (an < (1ll << 30) && den < (1ll << 24) && den >= 256) {  // |q| < 2^22: the estimate is off by at most 1-2
    int n32 = (int)num, d32 = (int)den;
    int q = (int)floorf(__fdividef((float)n32, (float)d32));
    int r = n32 - q * d32;
    while (r < 0) {
      q--;
      r += d32;
    }
    while (r >= d32) {
      q++;
      r -= d32;
    }
    return (long long)q;
  }
  if (an < (1ll << 52)) {
    long long q = (long long)floor((double)num / (double)den);
    long long r = num - q * den;
    while (r < 0) {
      q--;
      r += den;
    }
    while (r >= den) {
      q++;
      r -= den;
    }
    return q;
  }
  return floordiv64(num, den);
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

// m = the canvas CTM (Matrix2D order): 0.05 * shape matrix, entry by entry - ctx.scale(1/20, 1/20) followed by
// ctx.transform (canvas-renderer.ts:74, 179-188); Cairo transforms path points with it at path-build time.
__device__ __forceinline__ void to_device_fx(const double *m, double x, double y, int &fx, int &fy) {
  double px = m[0] * x;
  double qx = m[3] * y;
  double sx = px + qx;
  sx = sx + m[4];
  double py = m[2] * x;
  double qy = m[1] * y;
  double sy = py + qy;
  sy = sy + m[5];
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

// round-half-up of num / den for integer-valued doubles with |num| < 2^50, 0 < den < 2^25: every operand and
// the numerator 2 num + den are exact in FP64 and the quotient cannot round across an integer (the distance of a
// non-integer quotient to the next integer is >= 1/(2 den) >= 2^-26 while the rounding error is < 2^-27), so this
// is the exact integer floor((2 num + den) / (2 den)), i.e. the oracle's rdiv64.
// The division itself is replaced by a multiplication with inv2den = 1 / (2 den) (one FP64 division per segment
// instead of two per point): the estimate floor(N * inv2den) is off by at most one (relative error < 2^-51, quotient
// < 2^27), and the remainder N - q * 2 den is exact in FP64 (q * 2 den < 2^53), so one correction step gives the
// exact floor again.
__device__ __forceinline__ int rdiv_f64(double num, double den, double inv2den) {
  const double N = 2.0 * num + den, D = 2.0 * den;
  double q = floor(N * inv2den);
  const double r = N - q * D;
  if (r < 0.0)
    q -= 1.0;
  else if (r >= D)
    q += 1.0;
  return (int)q;
}

// den = n (lines) or n^2 (curves); inv2den = 1 / (2 den)
__device__ __forceinline__ double piece_inv2den(bool curve, int n) {
  return curve ? 1.0 / (2.0 * ((double)n * (double)n)) : 1.0 / (2.0 * (double)n);
}

__device__ __forceinline__ void piece_point(bool curve, const int *p, int n, int i, double inv2den, int &x, int &y) {
  if (!curve) {
    x = p[0] + rdiv_f64(((double)p[4] - (double)p[0]) * (double)i, (double)n, inv2den);
    y = p[1] + rdiv_f64(((double)p[5] - (double)p[1]) * (double)i, (double)n, inv2den);
  } else {
    // all products are exact: a, b, c <= n^2 <= 2^24 and |p| <= 2^24
    double a = (double)(n - i) * (double)(n - i), b = 2.0 * (double)i * (double)(n - i), c = (double)i * (double)i,
           nn = (double)n * (double)n;
    x = rdiv_f64(a * (double)p[0] + b * (double)p[2] + c * (double)p[4], nn, inv2den);
    y = rdiv_f64(a * (double)p[1] + b * (double)p[3] + c * (double)p[5], nn, inv2den);
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

// Fixed-point control points of segment `local` of a draw item (item fields are warp-uniform).
struct ItemRegs {
  double m[6];  // canvas CTM = 0.05 * Matrix2D
  uint32_t seg_first, path_off;
  uint32_t kind;  // ITEM_STATIC / ITEM_MORPH / ITEM_DYNAMIC
  double ratio;
};

__device__ __forceinline__ double item_ratio(const DrawItem &item) {
  const uint32_t kind = __ldg(&item.kind);
  return (kind & ITEM_RATIO_F32) ? (double)__ldg(&item.ratio_f32) : (double)__ldg(&item.ratio) / 65535.0;
}

__device__ __forceinline__ ItemRegs load_item(const RenderArgs &a, uint32_t it) {
  const DrawItem &item = a.items[it];
  ItemRegs r;
#pragma unroll
  for (int k = 0; k < 6; k++) r.m[k] = (double)__ldg(&item.m[k]) * 0.05;
  r.seg_first = __ldg(&item.seg_first);
  r.path_off = __ldg(&item.path_off);
  r.kind = __ldg(&item.kind) & ITEM_KIND_MASK;
  r.ratio = r.kind == ITEM_MORPH ? item_ratio(item) : 0.0;
  return r;
}

__device__ __forceinline__ void load_segment(const RenderArgs &a, const ItemRegs &item, uint32_t local, int p[6],
                                             bool &curve, uint32_t &pid) {
  double c[6];
  uint32_t pf;
  if (item.kind == ITEM_MORPH) {
    const SegMorph &s = a.segs_morph[item.seg_first + local];
#pragma unroll
    for (int k = 0; k < 6; k++) c[k] = lerp_ref((double)__ldg(&s.s[k]), (double)__ldg(&s.e[k]), item.ratio);
    pf = __ldg(&s.path_flags);
  } else {
    const SegStatic &s = (item.kind == ITEM_DYNAMIC ? a.segs_dynamic : a.segs_static)[item.seg_first + local];
#pragma unroll
    for (int k = 0; k < 6; k++) c[k] = (double)__ldg(&s.p[k]);
    pf = __ldg(&s.path_flags);
  }
  curve = (pf >> 31) != 0;
  pid = item.path_off + (pf & 0x7fffffffu);
  to_device_fx(item.m, c[0], c[1], p[0], p[1]);
  to_device_fx(item.m, c[4], c[5], p[4], p[5]);
  if (curve) {
    to_device_fx(item.m, c[2], c[3], p[2], p[3]);
  } else {
    p[2] = p[0];
    p[3] = p[1];
  }
}

// Depth chunks: items [chunk_first(c, f), chunk_first(c + 1, f)) of frame f belong to chunk c (pass-relative indices).
__device__ __forceinline__ uint32_t chunk_first(const RenderArgs &a, uint32_t c, uint32_t frame) {
  return __ldg(a.chunk_items + c * a.n_frames + frame);
}

// Unused entries of a stroke outline's reserved room in the dynamic segment store (k_stroke) are no segments at all.
__device__ __forceinline__ bool segment_is_null(const RenderArgs &a, const ItemRegs &item, uint32_t local) {
  return item.kind == ITEM_DYNAMIC && __ldg(&a.segs_dynamic[item.seg_first + local].path_flags) == kNullSegment;
}

// Path instance of segment `local` of a draw item (without transforming the segment).
__device__ __forceinline__ uint32_t segment_pid(const RenderArgs &a, const ItemRegs &item, uint32_t local) {
  uint32_t pf;
  if (item.kind == ITEM_MORPH)
    pf = __ldg(&a.segs_morph[item.seg_first + local].path_flags);
  else
    pf = __ldg(&(item.kind == ITEM_DYNAMIC ? a.segs_dynamic : a.segs_static)[item.seg_first + local].path_flags);
  return item.path_off + (pf & 0x7fffffffu);
}

__global__ void k_init(RenderArgs a) {
  pdl_enter();
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t stride = gridDim.x * blockDim.x;
  if (i == 0) {
    a.totals->n_edges = a.totals->n_slots = a.totals->n_records = 0;
    a.totals->overflow = 0;
    a.totals->error = 0;
    for (int k = 0; k < kMaxFineSlices; k++) a.totals->work[k] = 0;
    a.totals->n_list = 0;
    for (int k = 0; k < kMaxChunks; k++) a.totals->n_big_chunk[k] = a.totals->n_small_chunk[k] = a.totals->n_alive_items[k] = 0;
    a.totals->n_rowent = 0;
    a.totals->n_stage_blocks = 0;
    a.totals->overflow_stage = 0;
    a.totals->fine_hits = 0;
    a.totals->fine_records = 0;
    // the render before this one was aborted in this arena (working memory overflow): the self-cleaning arrays are
    // dirty until the host has zeroed them (recover), so this render stands down as well and is re-run with it
    if (*a.arena_dirty) a.totals->overflow = 32u;
  }
  for (uint32_t l = i; l < a.n_frames * (uint32_t)a.tiles_y; l += stride) a.row_count[l] = 0;
  for (uint32_t l = i; l < a.caps.stage / kStageBlock; l += stride) a.stage_used[l] = 0;
  for (uint32_t l = i; l < a.n_frames * (uint32_t)(a.tiles_x * a.tiles_y); l += stride) a.tile_cover[l] = 0;
  for (uint32_t l = i; l < a.n_frames * (uint32_t)a.tiles_y * a.cover_words; l += stride) a.cover_bits[l] = 0;
  for (uint32_t l = i; l < a.n_items; l += stride) a.item_alive[l] = 0;
}

// One warp per draw item, lanes over its segments (contiguous in the segment store).
// The ordered mode (one depth chunk): one warp per draw item, lanes over its (contiguous) segments - the draw item of
// every segment instance and the number of line pieces of every segment -> seg_edge_off, scanned into edge offsets.
// (With occlusion culling only visible segments are flattened, in no particular order, and k_flatten_emit<false>
// counts their pieces itself.)
__global__ void k_flatten_count(RenderArgs a) {
  pdl_enter();
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t it = warp; it < a.n_items; it += nwarps) {
    const uint32_t s0 = __ldg(a.item_seg_off + it), s1 = __ldg(a.item_seg_off + it + 1);
    if (s0 == s1) continue;
    const ItemRegs item = load_item(a, it);
    for (uint32_t j = s0 + lane; j < s1; j += 32) {
      int p[6];
      bool curve;
      uint32_t pid;
      uint32_t n = 0;
      if (!segment_is_null(a, item, j - s0)) {
        load_segment(a, item, j - s0, p, curve, pid);
        n = (uint32_t)piece_count(curve, p);
      }
      a.seg_edge_off[j] = n;
      a.seg_item[j] = it;
    }
  }
}

// Occlusion culling, per depth chunk (chunks are processed from the top one down).  tile_cover[tile] = 1 + the highest
// path instance found so far that covers the tile completely and opaquely (0 = none); every such path belongs to a
// chunk above the one being processed, so for the paths of this chunk "covered" is simply tile_cover != 0.  The same
// fact is kept as one bit per tile (cover_bits: cover_words 32-bit words per tile row, bit set = covered), which is
// what the visibility test of a whole path reads.
//
// k_path_alive, one THREAD per path instance of the chunk: the path is alive when at least one tile of its bbox is
// open, and its tile grid shrinks to the bounding box of its open tiles - what lies outside is clipped exactly like
// geometry outside the viewport (edges left of the grid still post their winding, see band_setup), so fewer slots are
// cleared, scanned and probed.  Dead paths are not flattened, not binned, get no records, and their slots are never
// read (k_fine checks path_alive first).  The lists of the chunk (draw items with a visible path for the flattener,
// visible paths with small / large grids for k_cover) are appended with one atomic per warp and list.  Nothing is
// cleared here: the per-slot record counters and winding deltas clean themselves (k_scatter counts the former back to
// zero, k_cover zeroes the latter as it reads them), so they are all zero whenever a chunk starts binning.
__global__ void __launch_bounds__(256) k_path_alive(RenderArgs a, uint32_t c) {
  pdl_enter();
  if (a.totals->overflow) return;
  const uint32_t frame = blockIdx.y;
  const uint32_t i0 = chunk_first(a, c, frame), i1 = chunk_first(a, c + 1, frame);
  // unordered edges: those of this chunk start at the current cursor (k_flatten_emit<false> of the chunk runs next)
  if (blockIdx.x == 0 && frame == 0 && threadIdx.x == 0) a.chunk_edge[c] = a.totals->n_edges;
  const uint32_t p0 = __ldg(a.item_path_off + i0), p1 = __ldg(a.item_path_off + i1);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t *bits = a.cover_bits + (size_t)frame * (size_t)a.tiles_y * a.cover_words;
  const bool top = c + 1 == a.n_chunks;  // nothing has been binned above the top chunk
  for (uint32_t base = p0 + blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < p1; base += stride) {
    const uint32_t pid = base + lane;
    bool alive = false;
    uint32_t n = 0, it = 0;
    int bx0 = 0, by0 = 0, bw = 0, bh = 0;
    int cmin = INT_MAX, cmax = -1, rmin = INT_MAX, rmax = -1;  // bounding box of the open tiles
    bool scan_small = false, scan_big = false;
    if (pid < p1) {
      const uint2 r = __ldg(reinterpret_cast<const uint2 *>(a.path_rec + pid));
      bx0 = r.x & 0xffff, by0 = r.x >> 16, bw = r.y & 0xffff, bh = r.y >> 16;
      // without culling (one chunk) every path is emitted, also those outside the viewport: the edge tap lists them all
      alive = bw > 0 || a.n_chunks == 1;
      if (alive && !top) {
        scan_small = bh <= 32 && ((bx0 + bw - 1) >> 5) - (bx0 >> 5) <= 1;  // up to 32 rows inside two 32-tile words
        scan_big = !scan_small;
      }
    }
    // The usual case, one thread per path: its rows one after the other, the two words of a row in flight together.  All
    // lanes run the trip count of the tallest grid of the warp (a grid of more than 32 rows, or wider than two words,
    // would keep the others waiting: those go to the warp-cooperative scan below).
    {
      const int hmax = __reduce_max_sync(0xffffffffu, scan_small ? bh : 0);
      const int w0 = bx0 >> 5;
      const bool two = ((bx0 + bw - 1) >> 5) != w0;
      const uint32_t in0 = scan_small ? ((0xffffffffu << (bx0 & 31)) & (two ? 0xffffffffu : (0xffffffffu >> (31 - ((bx0 + bw - 1) & 31))))) : 0u;
      const uint32_t in1 = (scan_small && two) ? (0xffffffffu >> (31 - ((bx0 + bw - 1) & 31))) : 0u;
      const uint32_t *row = bits + (size_t)by0 * a.cover_words + w0;
      uint32_t all0 = 0, all1 = 0;
      for (int y = 0; y < hmax; y++) {
        if (scan_small && y < bh) {
          const uint32_t o0 = ~__ldg(row) & in0;
          const uint32_t o1 = two ? (~__ldg(row + 1) & in1) : 0u;
          if (o0 | o1) {
            rmin = min(rmin, by0 + y);
            rmax = by0 + y;
          }
          all0 |= o0;
          all1 |= o1;
          row += a.cover_words;
        }
      }
      if (all0 | all1) {
        cmin = all0 ? w0 * 32 + __ffs(all0) - 1 : (w0 + 1) * 32 + __ffs(all1) - 1;
        cmax = all1 ? (w0 + 1) * 32 + 31 - __clz(all1) : w0 * 32 + 31 - __clz(all0);
      }
    }
    // large grids, one at a time with the whole warp: lanes over the rows (a 68-row grid scanned by its own thread
    // would keep the other 31 lanes waiting)
    for (uint32_t pending = __ballot_sync(0xffffffffu, scan_big); pending; pending &= pending - 1) {
      const int src = __ffs(pending) - 1;
      const int qx0 = __shfl_sync(0xffffffffu, bx0, src), qy0 = __shfl_sync(0xffffffffu, by0, src);
      const int qw = __shfl_sync(0xffffffffu, bw, src), qh = __shfl_sync(0xffffffffu, bh, src);
      const int w0 = qx0 >> 5, w1 = (qx0 + qw - 1) >> 5;
      int c0 = INT_MAX, c1 = -1, r0 = INT_MAX, r1 = -1;
      for (int y = qy0 + (int)lane; y < qy0 + qh; y += 32) {
        const uint32_t *row = bits + (size_t)y * a.cover_words;
        for (int w = w0; w <= w1; w++) {
          uint32_t open = ~__ldg(row + w);
          if (w == w0) open &= 0xffffffffu << (qx0 & 31);
          if (w == w1) open &= 0xffffffffu >> (31 - ((qx0 + qw - 1) & 31));
          if (open) {
            c0 = min(c0, w * 32 + __ffs(open) - 1);
            c1 = max(c1, w * 32 + 31 - __clz(open));
            r0 = min(r0, y);
            r1 = max(r1, y);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        c0 = min(c0, __shfl_xor_sync(0xffffffffu, c0, o));
        c1 = max(c1, __shfl_xor_sync(0xffffffffu, c1, o));
        r0 = min(r0, __shfl_xor_sync(0xffffffffu, r0, o));
        r1 = max(r1, __shfl_xor_sync(0xffffffffu, r1, o));
      }
      if ((int)lane == src) cmin = c0, cmax = c1, rmin = r0, rmax = r1;
    }
    if (pid < p1) {
      if (scan_small || scan_big) {
        alive = cmax >= 0;
        if (alive && (cmin != bx0 || rmin != by0 || cmax - cmin + 1 != bw || rmax - rmin + 1 != bh)) {
          bx0 = cmin, by0 = rmin, bw = cmax - cmin + 1, bh = rmax - rmin + 1;
          *reinterpret_cast<uint2 *>(a.path_rec + pid) = make_uint2((uint32_t)bx0 | ((uint32_t)by0 << 16), (uint32_t)bw | ((uint32_t)bh << 16));
        }
      }
      a.path_alive[pid] = alive ? 1u : 0u;
      a.path_rec_base[pid] = 0;
      if (alive) {
        n = (uint32_t)(bw * bh);
        it = a.path_item[pid];
      }
    }
    const uint32_t amask = __ballot_sync(0xffffffffu, alive);
    if (amask == 0) continue;
    // the draw items with something visible: k_flatten_emit<false> walks this list (one entry per item).  The paths of
    // an item are neighbours: one lane per distinct item of the warp claims it.
    bool claim = false;
    if (alive) {
      const uint32_t same = __match_any_sync(amask, it);
      if ((uint32_t)(__ffs(same) - 1) == lane) claim = atomicExch(&a.item_alive[it], 1u) == 0u;
    }
    // the chunk's visible paths for k_cover: small tile grids are scanned by one warp each, large ones by a block
    const bool big = alive && n > (uint32_t)kBackdropSmall, small = alive && n > 0 && !big;
    const uint32_t m_item = __ballot_sync(0xffffffffu, claim), m_big = __ballot_sync(0xffffffffu, big),
                   m_small = __ballot_sync(0xffffffffu, small);
    uint32_t b_item = 0, b_big = 0, b_small = 0;
    if (lane == 0) {
      if (m_item) b_item = atomicAdd(&a.totals->n_alive_items[c], (uint32_t)__popc(m_item));
      if (m_big) b_big = atomicAdd(&a.totals->n_big_chunk[c], (uint32_t)__popc(m_big));
      if (m_small) b_small = atomicAdd(&a.totals->n_small_chunk[c], (uint32_t)__popc(m_small));
    }
    b_item = __shfl_sync(0xffffffffu, b_item, 0);
    b_big = __shfl_sync(0xffffffffu, b_big, 0);
    b_small = __shfl_sync(0xffffffffu, b_small, 0);
    const uint32_t below = (1u << lane) - 1u;
    if (claim) a.alive_items[b_item + __popc(m_item & below)] = it;
    if (big) a.big_chunk[b_big + __popc(m_big & below)] = pid;
    if (small) a.small_chunk[b_small + __popc(m_small & below)] = pid;
  }
}

// Emits the flattened edges (16 bytes of geometry + the path instance index).  One warp takes 32 consecutive segment
// instances (their edges are contiguous in the output), and spreads the pieces of all of them evenly over its lanes:
// lane k computes the END point of piece k; the start point is the neighbour lane's end point, the segment's first
// control point, or the last point of the previous round.  Stores are coalesced 16-byte writes.
constexpr int kEmitWarps = 8;

// ORDERED (one depth chunk, every path emitted): piece counts and edge offsets come from k_flatten_count<true> + scan,
// so the edge list is in segment order (what the edge tap returns).  Otherwise the pieces are counted here, for
// visible segments only, and each warp takes the room for its 32 segments' edges from one cursor.
// 32 segment instances (lane `lane` holds segment j when `valid`; in the unordered mode all of them belong to draw
// item it_known): pieces of the visible ones, spread evenly over the lanes.
template <bool ORDERED>
__device__ __forceinline__ void emit_segments(const RenderArgs &a, int (*sh_p)[6], double *sh_inv, uint32_t *sh_pid, uint32_t lane,
                                              uint32_t j, bool valid, uint32_t it_known) {
    int n = 0;
    uint32_t off = 0;
    if (valid) {
      const uint32_t it = ORDERED ? a.seg_item[j] : it_known;
      bool visible = true;
      uint32_t local = 0;
      {
        local = j - __ldg(a.item_seg_off + it);
        ItemRegs head;  // the three fields segment_pid needs; the matrix is loaded for visible paths only
        head.seg_first = __ldg(&a.items[it].seg_first);
        head.path_off = __ldg(&a.items[it].path_off);
        head.kind = __ldg(&a.items[it].kind) & ITEM_KIND_MASK;
        visible = !segment_is_null(a, head, local) && __ldg(a.path_alive + segment_pid(a, head, local)) != 0;
      }
      if (visible) {  // hidden paths emit nothing (their edges stay marked ~0)
        const ItemRegs item = load_item(a, it);
        int p[6];
        bool curve;
        uint32_t pid;
        load_segment(a, item, local, p, curve, pid);
        if (ORDERED) {
          off = a.seg_edge_off[j];
          n = (int)(a.seg_edge_off[j + 1] - off);
        } else {
          n = piece_count(curve, p);
        }
#pragma unroll
        for (int k = 0; k < 6; k++) sh_p[lane][k] = p[k];
        sh_inv[lane] = piece_inv2den(curve, n);
        sh_pid[lane] = pid | (curve ? 0x80000000u : 0u);
      }
    }
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - n;
    bool room = true;
    if (!ORDERED) {
      uint32_t base_e = 0;
      if (lane == 0 && total) base_e = atomicAdd(&a.totals->n_edges, (uint32_t)total);
      base_e = __shfl_sync(0xffffffffu, base_e, 0);
      off = base_e + (uint32_t)excl;
      if (total && (base_e + (uint32_t)total > a.caps.edges || base_e + (uint32_t)total < base_e)) {
        if (lane == 0) atomicOr(&a.totals->overflow, 1u);
        room = false;
      }
    }
    __syncwarp();
    int carry_x = 0, carry_y = 0;
    for (int k0 = 0; k0 < total; k0 += 32) {
      const int k = min(k0 + (int)lane, total - 1);
      int o = 0;  // owner = the last lane whose first piece is <= k
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        int v = __shfl_sync(0xffffffffu, excl, o + step);
        if (v <= k) o += step;
      }
      const int first = __shfl_sync(0xffffffffu, excl, o);
      const int n_o = __shfl_sync(0xffffffffu, n, o);
      const uint32_t off_o = __shfl_sync(0xffffffffu, off, o);  // first edge of the owner's segment
      const int i = k - first + 1;  // 1..n_o
      int p[6];
#pragma unroll
      for (int q = 0; q < 6; q++) p[q] = sh_p[o][q];
      const uint32_t pc = sh_pid[o];
      int qx, qy;
      piece_point((pc >> 31) != 0, p, n_o, i, sh_inv[o], qx, qy);
      int px = __shfl_up_sync(0xffffffffu, qx, 1), py = __shfl_up_sync(0xffffffffu, qy, 1);
      if (i == 1) {
        px = p[0];
        py = p[1];
      } else if (lane == 0) {
        px = carry_x;
        py = carry_y;
      }
      carry_x = __shfl_sync(0xffffffffu, qx, 31);
      carry_y = __shfl_sync(0xffffffffu, qy, 31);
      if (room && k0 + (int)lane < total) {
        a.edges[off_o + (uint32_t)(i - 1)] = make_int4(px, py, qx, qy);
        a.edge_pid[off_o + (uint32_t)(i - 1)] = pc & 0x7fffffffu;
      }
    }
    __syncwarp();
}

template <bool ORDERED>
__global__ void __launch_bounds__(kEmitWarps * 32) k_flatten_emit(RenderArgs a, uint32_t c) {
  pdl_enter();
  if (a.totals->overflow) return;
  __shared__ int sh_p[kEmitWarps][32][6];
  __shared__ double sh_inv[kEmitWarps][32];
  __shared__ uint32_t sh_pid[kEmitWarps][32];  // path instance | curve << 31
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (ORDERED) {
    // segment instances of depth chunk c in frame blockIdx.y, in order
    const uint32_t stride = gridDim.x * kEmitWarps * 32;
    const uint32_t s_begin = __ldg(a.item_seg_off + chunk_first(a, c, blockIdx.y));
    const uint32_t s_end = __ldg(a.item_seg_off + chunk_first(a, c + 1, blockIdx.y));
    for (uint32_t base = s_begin + (blockIdx.x * kEmitWarps + w) * 32; base < s_end; base += stride)
      emit_segments<true>(a, sh_p[w], sh_inv[w], sh_pid[w], lane, base + lane, base + lane < s_end, 0u);
  } else {
    // the draw items of the chunk with a visible path (listed by k_path_alive), one warp per item
    const uint32_t n_alive = a.totals->n_alive_items[c];
    const uint32_t nwarps = gridDim.x * gridDim.y * kEmitWarps;
    for (uint32_t ai = (blockIdx.y * gridDim.x + blockIdx.x) * kEmitWarps + w; ai < n_alive; ai += nwarps) {
      const uint32_t it = a.alive_items[ai];
      const uint32_t s0 = __ldg(a.item_seg_off + it), s1 = __ldg(a.item_seg_off + it + 1);
      for (uint32_t base = s0; base < s1; base += 32) emit_segments<false>(a, sh_p[w], sh_inv[w], sh_pid[w], lane, base + lane, base + lane < s1, it);
    }
  }
}

// ======================================================================================================
// Device stroker: the outlines of morph-shape strokes, per draw, at the draw's ratio
// (canvas-renderer.ts:252-266: lerped path and width, round caps and joins).  One thread per draw runs the streaming
// generator of stroke_core.h - FP64 + - * / sqrt in the host stroker's order, so the float32 segments are the oracle's
// bit for bit - over the morph shape's lines, writes the segments into the room reserved for the draw in the dynamic
// segment store, the bounds of every visible line path into its dynamic paint, and marks the unused room empty.  An
// outline that outgrows its room raises overflow bit 6: the host lays the batch out again with exact counts.
// ======================================================================================================
__global__ void __launch_bounds__(64) k_stroke(RenderArgs a) {
  pdl_enter();
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.n_jobs) return;
  const StrokeJob job = a.jobs[j];
  SegStatic *out = a.segs_dynamic + job.seg_first;
  double width_state = 1.0;  // Canvas default lineWidth; a zero width is ignored and the previous one stays
  uint32_t used = 0, need = 0, path = 0;
  for (uint32_t l = 0; l < job.line_count; l++) {
    const stroke::LineDev ln = a.mlines[job.line_first + l];
    const double w = stroke::lerp(ln.w0, ln.w1, job.ratio);
    if (w > 0) width_state = w;
    const double al = stroke::lerp(ln.color0[3] / 255.0, ln.color1[3] / 255.0, job.ratio);
    if (al <= 0) continue;  // composites nothing (the host left it out of the draw's paths the same way)
    stroke::Sink sink{out + used, job.seg_cap - used, 0, path, {1.f, 1.f, 0.f, 0.f}, false, 0.f, 0.f, 0.f, 0.f};
    stroke::stroke_line(a.mcmds + ln.cmd_first, ln.cmd_count, job.ratio, width_state, sink);
    need += sink.n;
    used += min(sink.n, job.seg_cap - used);
    if (path < job.path_count) {
      float *b = a.paints_dynamic[job.paint_first + path].bounds;
      b[0] = sink.bounds[0], b[1] = sink.bounds[1], b[2] = sink.bounds[2], b[3] = sink.bounds[3];
    }
    path++;
  }
  for (uint32_t k = used; k < job.seg_cap; k++) out[k].path_flags = kNullSegment;
  if (need > job.seg_cap) atomicOr(&a.totals->overflow, 64u);
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

// n = n_ptr ? min(*n_ptr, n_cap) : n_cap.  src and dst may alias (in-place scan).
__global__ void k_scan_partials(const uint32_t *__restrict__ src, const uint32_t *n_ptr, uint32_t n_cap,
                                uint32_t *partials) {
  pdl_enter();
  __shared__ uint32_t sh[32];
  uint32_t n = n_ptr ? min(*n_ptr, n_cap) : n_cap;
  uint32_t chunk = scan_chunk(n);
  uint64_t begin = (uint64_t)blockIdx.x * chunk;
  uint64_t end = min((uint64_t)n, begin + chunk);
  uint32_t s = 0;
  for (uint64_t i = begin + threadIdx.x; i < end; i += blockDim.x) s += src[i];
  s = block_reduce(s, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// exclusive block scan of one value per thread (blockDim.x <= 1024); returns the exclusive prefix, *total = sum
__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, uint32_t *sh, uint32_t *total) {
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if ((threadIdx.x & 31) >= o) inc += t;
  }
  if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = inc;
  __syncthreads();
  uint32_t wbase = 0, tot = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
    uint32_t x = sh[w];
    if (w < (int)(threadIdx.x >> 5)) wbase += x;
    tot += x;
  }
  __syncthreads();
  *total = tot;
  return wbase + inc - v;
}

// Scans the block partials (one thread each); stores the grand total at dst[n] and *total_out; raises
// overflow_bit when the total exceeds total_cap.
__global__ void __launch_bounds__(1024) k_scan_spine(uint32_t *partials, uint32_t *dst, const uint32_t *n_ptr,
                                                    uint32_t n_cap, uint32_t *total_out, uint32_t total_cap,
                                                    uint32_t *overflow, uint32_t overflow_bit) {
  pdl_enter();
  __shared__ uint32_t sh[32];
  uint32_t n = n_ptr ? min(*n_ptr, n_cap) : n_cap;
  uint32_t v = threadIdx.x < (uint32_t)kScanBlocks ? partials[threadIdx.x] : 0u;
  uint32_t total;
  uint32_t ex = block_exclusive(v, sh, &total);
  if (threadIdx.x < (uint32_t)kScanBlocks) partials[threadIdx.x] = ex;
  if (threadIdx.x == 0) {
    dst[n] = total;
    if (total_out) *total_out = total;
    if (overflow && total > total_cap) atomicOr(overflow, overflow_bit);
  }
}

__global__ void k_scan_apply(const uint32_t *src, uint32_t *dst, const uint32_t *n_ptr, uint32_t n_cap,
                             const uint32_t *partials) {
  pdl_enter();
  __shared__ uint32_t sh[32];
  uint32_t n = n_ptr ? min(*n_ptr, n_cap) : n_cap;
  uint32_t chunk = scan_chunk(n);
  uint64_t begin = (uint64_t)blockIdx.x * chunk;
  uint64_t end = min((uint64_t)n, begin + chunk);
  uint32_t carry = partials[blockIdx.x];
  for (uint64_t base = begin; base < end; base += kScanTile) {
    uint64_t i0 = base + (uint64_t)threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (i0 + k < end) ? src[i0 + k] : 0u;
    uint32_t tile_total;
    uint32_t ex = carry + block_exclusive(v[0] + v[1] + v[2] + v[3], sh, &tile_total);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (i0 + k < end) dst[i0 + k] = ex;
      ex += v[k];
    }
    carry += tile_total;
  }
}

// Whole scan in one block, for short arrays (n <= kScanSmallMax, known on the host).
constexpr uint32_t kScanSmallMax = 16384;
__global__ void __launch_bounds__(1024) k_scan_small(const uint32_t *src, uint32_t *dst, uint32_t n, uint32_t *total_out,
                                                    uint32_t total_cap, uint32_t *overflow, uint32_t overflow_bit) {
  pdl_enter();
  __shared__ uint32_t sh[32];
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n; base += 4096) {
    uint32_t i0 = base + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (i0 + k < n) ? src[i0 + k] : 0u;
    uint32_t tile_total;
    uint32_t ex = carry + block_exclusive(v[0] + v[1] + v[2] + v[3], sh, &tile_total);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (i0 + k < n) dst[i0 + k] = ex;
      ex += v[k];
    }
    carry += tile_total;
  }
  if (threadIdx.x == 0) {
    dst[n] = carry;
    if (total_out) *total_out = carry;
    if (overflow && carry > total_cap) atomicOr(overflow, overflow_bit);
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

// swf-tree ColorTransformWithAlpha (include/swfr.h swfr_color_transform; oracle/raster.c cx_channel / cx_premul / cx_solid)
__device__ __forceinline__ int cx_channel(int c, int mult, int add) {
  const int v = ((c * mult) >> 8) + add;  // |c * mult| < 2^23: the arithmetic shift is the oracle's floor
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}
__device__ uint32_t cx_premul(uint32_t p, const int16_t *cx) {
  const int a = (int)(p >> 24);
  const int a2 = cx_channel(a, cx[3], cx[7]);
  uint32_t out = (uint32_t)a2 << 24;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int c = (int)((p >> (8 * k)) & 255u);
    int s = a ? (c * 255 + a / 2) / a : 0;
    if (s > 255) s = 255;
    const int c2 = cx_channel(s, cx[k], cx[4 + k]);
    out |= (uint32_t)((c2 * a2 + 127) / 255) << (8 * k);
  }
  return out;
}
__device__ uint32_t cx_solid(double r8, double g8, double b8, double alpha, const int16_t *cx) {
  double af = (double)(float)alpha;
  if (af < 0) af = 0;
  if (af > 1) af = 1;
  const int a8 = (int)(af * 255.0 + 0.5);
  const int r2 = cx_channel((int)r8, cx[0], cx[4]), g2 = cx_channel((int)g8, cx[1], cx[5]), b2 = cx_channel((int)b8, cx[2], cx[6]);
  const int a2 = cx_channel(a8, cx[3], cx[7]);
  return solid_premul8((double)r2, (double)g2, (double)b2, a2 / 255.0);
}

__device__ uint32_t morph_solid(const uint8_t *c0, const uint8_t *c1, double r, const int16_t *cx) {
  double ch[4];
  for (int i = 0; i < 4; i++) ch[i] = lerp_ref(c0[i] / 255.0, c1[i] / 255.0, r);
  double red = (double)(((long long)(ch[0] * 255.0)) & 0xff);
  double g = ceil((double)(float)(ch[1] * 255.0));
  double b = ceil((double)(float)(ch[2] * 255.0));
  if (g < 0) g = 0;
  if (g > 255) g = 255;
  if (b < 0) b = 0;
  if (b > 255) b = 255;
  return cx ? cx_solid(red, g, b, ch[3], cx) : solid_premul8(red, g, b, ch[3]);
}

constexpr int kMaxTileRows = 1024;  // frames up to 16384 px high

__global__ void __launch_bounds__(256) k_path_setup(RenderArgs a) {
  pdl_enter();
  // candidate lists, step 0: how many path instances touch each tile row of the frame.  Counted in shared memory
  // for the frame of the block's first path (a block of consecutive paths rarely spans two frames; the others go
  // straight to global memory), then flushed with one atomic per non-empty row.
  __shared__ uint32_t sh_rows[kMaxTileRows];
  uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t pid0 = blockIdx.x * blockDim.x; pid0 < a.n_paths; pid0 += stride) {
  for (int y = threadIdx.x; y < a.tiles_y; y += blockDim.x) sh_rows[y] = 0;
  const uint32_t block_frame = a.items[find_owner(a.item_path_off, a.n_items, pid0)].frame;
  __syncthreads();
  const uint32_t pid = pid0 + threadIdx.x;
  if (pid < a.n_paths) {
    uint32_t it = find_owner(a.item_path_off, a.n_items, pid);
    const DrawItem &item = a.items[it];
    const DefPaint &dp = ((item.kind & ITEM_KIND_MASK) == ITEM_DYNAMIC ? a.paints_dynamic
                                                                       : a.def_paints)[item.paint_first + (pid - a.item_path_off[it])];
    // ---- tile bbox: the device-space hull of the corners of the path's control-point bounds (both morph states).
    // The transform is affine, so every control point of the instance lies inside it up to rounding; one 1/256 px
    // unit of padding covers that.  (Exact for translations and scales, looser under rotation.)
    int minx = 1, miny = 1, maxx = 0, maxy = 0;
    if (dp.bounds[0] <= dp.bounds[2]) {
      double ctm[6];
#pragma unroll
      for (int k = 0; k < 6; k++) ctm[k] = (double)item.m[k] * 0.05;
      minx = miny = INT_MAX;
      maxx = maxy = INT_MIN;
      // The bounds hold both morph states.  A ratio outside [0, 1] (the reference lerps any number,
      // canvas-renderer.ts:24-26) extrapolates every coordinate by at most `over` times the extent of the bounds:
      // the box grows by that much, so the geometry is not cut off at the bbox.
      double bnd[4] = {(double)dp.bounds[0], (double)dp.bounds[1], (double)dp.bounds[2], (double)dp.bounds[3]};
      if ((item.kind & ITEM_KIND_MASK) == ITEM_MORPH) {
        const double rr = item_ratio(item);
        const double over = rr < 0.0 ? -rr : (rr > 1.0 ? rr - 1.0 : 0.0);
        if (over > 0.0) {
          const double px = over * (bnd[2] - bnd[0]), py = over * (bnd[3] - bnd[1]);
          bnd[0] -= px, bnd[2] += px, bnd[1] -= py, bnd[3] += py;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        int fx, fy;
        to_device_fx(ctm, bnd[(k & 1) ? 2 : 0], bnd[(k & 2) ? 3 : 1], fx, fy);
        minx = min(minx, fx), maxx = max(maxx, fx);
        miny = min(miny, fy), maxy = max(maxy, fy);
      }
      minx -= 1, miny -= 1, maxx += 1, maxy += 1;
    }
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
    double ratio = item_ratio(item);
    const int16_t *cx = (item.kind & ITEM_CX) ? a.item_cx + (size_t)it * 8 : nullptr;
    if (dp.type == PAINT_SOLID) {
      if (dp.flags & PF_COLOR_MORPH)
        rec.color = morph_solid(dp.color0, dp.color1, ratio, cx);
      else if (cx)
        rec.color = cx_solid(dp.color0[0], dp.color0[1], dp.color0[2], dp.color0[3] / 255.0, cx);
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
            pi.focal = 1.0f / pi.rx;  // bitmaps: focal / omf carry 1 / rx, 1 / ry
            pi.omf = 1.0f / pi.ry;
            pi.inv_bw = 1.0f / (float)bm.w;
            pi.inv_bh = 1.0f / (float)bm.h;
            if (bm.opaque && dp.repeating) flags |= 1u;
          }
        } else {
          double fp = dp.type == PAINT_FOCAL ? dp.focal : 0.0;
          if (fp > 0.98) fp = 0.98;
          if (fp < -0.98) fp = -0.98;
          pi.focal = (float)fp;
          pi.omf = (float)(1.0 - fp * fp);
          pi.rx = 1.0f / pi.omf;  // gradients: rx carries 1 / omf
          pi.ptr = (unsigned long long)(a.ramps + (size_t)dp.lut * kRampSize);
          if (dp.flags & PF_OPAQUE_RAMP) flags |= 1u;
        }
      }
      if (cx) {
        // every evaluated pixel goes through the transform; an opaque paint stays opaque only if alpha 255 maps to 255
        pi.cx_on = 1;
#pragma unroll
        for (int k = 0; k < 8; k++) pi.cx[k] = cx[k];
        if (cx_channel(255, cx[3], cx[7]) != 255) flags &= ~1u;
      }
      a.paint_inst[pid] = pi;
    }
    if (dp.flags & PF_SAMPLED) flags |= 2u;  // a stroke outline: k_fine applies the non-zero rule per sub-scanline
    if (!valid) bw = bh = 0;
    rec.xy0 = (uint32_t)bx0 | ((uint32_t)by0 << 16);
    rec.wh = (uint32_t)bw | ((uint32_t)bh << 16);
    rec.info = dp.type | (flags << 8) | (item.frame << 16);  // frames per pass < 65536
    a.path_rec[pid] = rec;
    a.path_slot_off[pid] = (uint32_t)(bw * bh);
    a.path_item[pid] = it;
    if (bw > 0) {
      if (item.frame == block_frame) {
        for (int y = 0; y < bh; y++) atomicAdd(&sh_rows[by0 + y], 1u);
      } else {
        uint32_t *rc = a.row_count + item.frame * (uint32_t)a.tiles_y + (uint32_t)by0;
        for (int y = 0; y < bh; y++) atomicAdd(rc + y, 1u);
      }
    }
  }
  __syncthreads();
  for (int y = threadIdx.x; y < a.tiles_y; y += blockDim.x)
    if (sh_rows[y]) atomicAdd(a.row_count + block_frame * (uint32_t)a.tiles_y + (uint32_t)y, sh_rows[y]);
  __syncthreads();
  }
}

// Candidate lists: for every (frame, tile row, group of kGroupTiles tile columns) the path instances whose tile bbox
// overlaps it, IN PAINT ORDER.  Built without atomics on the data path and without sorting, by two ordered
// compactions (ballot + popc prefix):
//   k_row_lists   - one block per (frame, tile row) sweeps the frame's path instances in order and keeps those whose
//                   bbox covers the row, as (path instance, tile x-range); it also counts entries per column group;
//   k_group_lists - one warp per (frame, row, group) sweeps that row list in order and keeps the overlapping ones.
// The visible path instances of every frame, in paint order (ordered compaction, one block per frame): what the row
// lists are built from.
__global__ void __launch_bounds__(1024) k_alive_paths(RenderArgs a) {
  pdl_enter();
  if (a.totals->overflow) return;
  __shared__ uint32_t sh_warp[32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (uint32_t frame = blockIdx.x; frame < a.n_frames; frame += gridDim.x) {
    const uint32_t p0 = a.frame_path_off[frame], p1 = a.frame_path_off[frame + 1];
    uint32_t out = p0;
    for (uint32_t base = p0; base < p1; base += 1024) {
      const uint32_t pid = base + tid;
      const bool hit = pid < p1 && __ldg(a.path_alive + pid) != 0 && (__ldg(&a.path_rec[pid].wh) & 0xffffu) != 0;
      const uint32_t mask = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) sh_warp[w] = __popc(mask);
      __syncthreads();
      uint32_t wbase = 0, tot = 0;
#pragma unroll
      for (int k = 0; k < 32; k++) {
        const uint32_t c = sh_warp[k];
        if (k < (int)w) wbase += c;
        tot += c;
      }
      if (hit) a.alive_paths[out + wbase + __popc(mask & ((1u << lane) - 1u))] = pid;
      out += tot;
      __syncthreads();
    }
    if (tid == 0) a.alive_count[frame] = out - p0;
    __syncthreads();
  }
}

constexpr int kRowThreads = 256;
constexpr int kMaxGroups = 512;  // tile columns <= 4096: frames up to 65536 px wide

__global__ void __launch_bounds__(kRowThreads) k_row_lists(RenderArgs a) {
  pdl_enter();
  if (a.totals->overflow) return;
  __shared__ uint32_t sh_warp[kRowThreads / 32];
  __shared__ uint32_t sh_grp[kMaxGroups];
  const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t n_rows = a.n_frames * (uint32_t)a.tiles_y;
  for (uint32_t rl = blockIdx.x; rl < n_rows; rl += gridDim.x) {
    const uint32_t frame = rl / (uint32_t)a.tiles_y;
    const int row = (int)(rl - frame * (uint32_t)a.tiles_y);
    // the frame's visible paths, in paint order (k_alive_paths): hidden ones are left out of the lists
    const uint32_t p0 = a.frame_path_off[frame], p1 = p0 + a.alive_count[frame];
    for (uint32_t g = tid; g < a.groups_x; g += kRowThreads) sh_grp[g] = 0;
    __syncthreads();
    uint32_t out = a.row_off[rl];
    for (uint32_t base = p0; base < p1; base += kRowThreads) {
      bool hit = false;
      uint32_t xr = 0, pid = 0;
      if (base + tid < p1) {
        pid = __ldg(a.alive_paths + base + tid);
        const uint2 r = __ldg(reinterpret_cast<const uint2 *>(a.path_rec + pid));  // xy0, wh
        const int by0 = r.x >> 16, bw = r.y & 0xffff, bh = r.y >> 16;
        hit = row >= by0 && row < by0 + bh;
        xr = (r.x & 0xffffu) | ((uint32_t)bw << 16);
      }
      const uint32_t mask = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) sh_warp[w] = __popc(mask);
      __syncthreads();
      uint32_t wbase = 0, tot = 0;
#pragma unroll
      for (int k = 0; k < kRowThreads / 32; k++) {
        const uint32_t c = sh_warp[k];
        if (k < (int)w) wbase += c;
        tot += c;
      }
      if (hit) {
        a.row_items[out + wbase + __popc(mask & ((1u << lane) - 1u))] = make_uint2(pid, xr);
        const uint32_t bx0 = xr & 0xffffu, bw = xr >> 16;
        const uint32_t g0 = bx0 / kGroupTiles, g1 = (bx0 + bw - 1) / kGroupTiles;
        for (uint32_t g = g0; g <= g1; g++) atomicAdd(&sh_grp[g], 1u);
      }
      out += tot;
      __syncthreads();
    }
    for (uint32_t g = tid; g < a.groups_x; g += kRowThreads) a.list_off[rl * a.groups_x + g] = sh_grp[g];
    if (tid == 0) a.row_count[rl] = out - a.row_off[rl];  // entries kept (the offsets were sized for all paths)
    __syncthreads();
  }
}

__global__ void k_group_lists(RenderArgs a) {
  pdl_enter();
  if (a.totals->overflow) return;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t l = warp; l < a.n_lists; l += nwarps) {
    const uint32_t rl = l / a.groups_x, g = l - rl * a.groups_x;
    const uint32_t r0 = a.row_off[rl], r1 = r0 + a.row_count[rl];
    uint32_t out = a.list_off[l];
    const uint32_t gx0 = g * kGroupTiles, gx1 = gx0 + kGroupTiles;
    for (uint32_t base = r0; base < r1; base += 32) {
      const uint32_t i = base + lane;
      bool hit = false;
      uint32_t pid = 0;
      if (i < r1) {
        const uint2 e = __ldg(a.row_items + i);
        const uint32_t bx0 = e.y & 0xffffu, bw = e.y >> 16;
        hit = bx0 < gx1 && bx0 + bw > gx0;
        pid = e.x;
      }
      const uint32_t mask = __ballot_sync(0xffffffffu, hit);
      if (hit) a.list_items[out + __popc(mask & ((1u << lane) - 1u))] = pid;
      out += __popc(mask);
    }
  }
}


// ======================================================================================================
// K2: tile binning (count, then scatter of tile-clipped 8-byte records)
// ======================================================================================================

// round-half-up of p * q / d (d != 0).  SMALL: |p|, |q|, |d| <= 2^14 (an edge of at most 64 px per axis) and
// |p| <= |d|, so everything fits 32 bits and |result| <= 2^14: a float quotient is then off by at most one and a
// single remainder check makes it the exact integer (same value as the oracle's rdiv64).
template <bool SMALL>
__device__ __forceinline__ int muldiv(int p, int q, int d) {
  if (SMALL) {
    int a = p * q;
    if (d < 0) {
      a = -a;
      d = -d;
    }
    int num = 2 * a + d, den = 2 * d;
    int qq = (int)floorf(__fdividef((float)num, (float)den));
    int r = num - qq * den;
    qq += (int)(r >= den) - (int)(r < 0);
    return qq;
  }
  return (int)rdiv64((long long)p * (long long)q, (long long)d);
}

__device__ __forceinline__ unsigned long long pack_record(int xa, int ya, int xb, int yb, int fs, int fe) {
  return (unsigned long long)xa | ((unsigned long long)ya << 13) | ((unsigned long long)xb << 26) |
         ((unsigned long long)yb << 39) | ((unsigned long long)fs << 52) | ((unsigned long long)fe << 53);
}

// The part of one edge that lies in tile row (band) `b` of its path's grid: the band-clipped sub-edge, the slots of
// the band's row and the range of tile columns it touches.  The crossings with the band's top and bottom lines come
// straight from the edge's end points (x0 + round((Y - y0) (x1 - x0) / (y1 - y0))), so every band of an edge can be
// processed independently of the others.  Also posts the backdrop deltas of the band.
struct BandPiece {
  int xs, ys, xe, ye;  // sub-edge, in edge direction
  int c0, c1;          // tile columns touched inside the grid (c0 > c1: none)
  uint32_t row_base;   // slot of column bx0 in this band
  int Yt;
};

template <bool SMALL>
__device__ __forceinline__ BandPiece band_setup(const RenderArgs &a, int x0, int y0, int x1, int y1, int bx0, int by0,
                                                int bw, uint32_t slot_base, int b) {
  const int B = kTileFx;
  BandPiece p;
  const bool horiz = y0 == y1;
  const bool down = y0 < y1;
  const int ylo = min(y0, y1), yhi = max(y0, y1);
  const int Yt = b * B, Yb = Yt + B;
  if (horiz) {
    p.xs = x0, p.ys = y0, p.xe = x1, p.ye = y1;
  } else {
    const int yu = max(ylo, Yt), yl = min(yhi, Yb);
    const int xu = (ylo >= Yt) ? (down ? x0 : x1) : x0 + muldiv<SMALL>(Yt - y0, x1 - x0, y1 - y0);
    const int xl = (yhi <= Yb) ? (down ? x1 : x0) : x0 + muldiv<SMALL>(Yb - y0, x1 - x0, y1 - y0);
    if (down) {
      p.xs = xu, p.ys = yu, p.xe = xl, p.ye = yl;
    } else {
      p.xs = xl, p.ys = yl, p.xe = xu, p.ye = yu;
    }
  }
  p.Yt = Yt;
  p.row_base = slot_base + (uint32_t)((b - by0) * bw);
  if (p.ys == Yt) {
    int lx = max((p.xs >> 12) + 1 - bx0, 0);
    if (lx < bw) atomicAdd(&a.slot_backdrop[p.row_base + lx], 1);
  }
  if (p.ye == Yt) {
    int lx = max((p.xe >> 12) + 1 - bx0, 0);
    if (lx < bw) atomicAdd(&a.slot_backdrop[p.row_base + lx], -1);
  }
  const int xlo = min(p.xs, p.xe), xhi = max(p.xs, p.xe);
  p.c0 = max(xlo >> 12, bx0);
  p.c1 = min(xhi >> 12, bx0 + bw - 1);
  return p;
}


// One tile column `t` of a band piece: the tile-clipped record (or an invalid marker when the clipped piece is a
// point).  The crossings with the tile's left and right boundary lines come straight from the band piece
// (ys + round((X - xs) (ye - ys) / (xe - xs))), so every column can be processed independently of the others.
template <bool SMALL>
__device__ __forceinline__ bool column_record(int xs, int ys, int xe, int ye, int Yt, int t, unsigned long long &rc) {
  const int B = kTileFx;
  const int xlo = min(xs, xe), xhi = max(xs, xe);
  const bool right = xs < xe;
  const int X0 = t * B, X1 = X0 + B;
  const bool clip_l = xlo < X0, clip_r = xhi > X1;
  int y_left = 0, y_right = 0;
  if (clip_l) y_left = ys + muldiv<SMALL>(X0 - xs, ye - ys, xe - xs);
  if (clip_r) y_right = ys + muldiv<SMALL>(X1 - xs, ye - ys, xe - xs);
  int ax, ay, bx, by, fs = 0, fe = 0;
  if (right) {
    ax = clip_l ? X0 : xs, ay = clip_l ? y_left : ys, fs = clip_l;
    bx = clip_r ? X1 : xe, by = clip_r ? y_right : ye;
  } else {
    ax = clip_r ? X1 : xs, ay = clip_r ? y_right : ys;
    bx = clip_l ? X0 : xe, by = clip_l ? y_left : ye, fe = clip_l;
  }
  rc = pack_record(ax - X0, ay - Yt, bx - X0, by - Yt, fs, fe);
  return !(ay == by && !fs && !fe);
}

// K2, pass 1 of 2.  Two levels of warp-cooperative expansion keep all lanes busy although edges touch different
// numbers of tiles:
//   level 1 - a warp takes 32 consecutive edges, counts the bands (tile rows) each of them touches inside its path's
//             grid and spreads the (edge, band) pairs evenly over its lanes (warp prefix sum + search); every lane
//             clips its band piece, posts the backdrop deltas and counts the tile columns the piece touches;
//   level 2 - the (piece, column) pairs of the round are spread over the lanes the same way; every lane clips ONE
//             record, counts it in its slot and writes it to the warp's staging block (consecutive lanes write
//             consecutive 16-byte entries; one global atomic per kStageBlock entries).
// Pass 2 (k_scatter) moves the staged records to their slots once the offsets are known, without touching the
// geometry again.
constexpr int kBinWarps = 8;

template <bool ORDERED>
__global__ void __launch_bounds__(kBinWarps * 32) k_bin(RenderArgs a, uint32_t c) {
  pdl_enter();
  if (a.totals->overflow) return;
  __shared__ uint32_t sh_cover[kBinWarps][32];  // per edge: first tile of its frame in the cover map
  __shared__ int4 sh_edge[kBinWarps][32];
  __shared__ uint4 sh_path[kBinWarps][32];   // xy0, bw | first band << 16, slot base, path instance
  __shared__ int4 sh_piece[kBinWarps][32];   // band piece xs, ys, xe, ye
  __shared__ uint4 sh_pmeta[kBinWarps][32];  // first column, slot of that column, band row, path instance | small << 31
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t stride = gridDim.x * kBinWarps * 32;
  const uint32_t cap_blocks = a.caps.stage / kStageBlock;
  // ORDERED: the edges of depth chunk c in frame blockIdx.y (segments, hence edges, are stored in item order);
  // otherwise the edges the chunk's k_flatten_emit appended, of all frames (grid.y = 1)
  uint32_t e_begin, e_end;
  if (ORDERED) {
    e_begin = a.seg_edge_off[__ldg(a.item_seg_off + chunk_first(a, c, blockIdx.y))];
    e_end = a.seg_edge_off[__ldg(a.item_seg_off + chunk_first(a, c + 1, blockIdx.y))];
  } else {
    e_begin = a.chunk_edge[c];
    e_end = min(a.totals->n_edges, a.caps.edges);
  }
  const uint32_t tiles = (uint32_t)(a.tiles_x * a.tiles_y);
  // staging blocks: a thread block that has edges takes one block per warp with a single atomic (19 000 warps asking
  // one counter each is what the kernel used to wait for); a warp that fills its block takes further ones itself
  __shared__ uint32_t sh_first_blk;
  if (threadIdx.x == 0 && e_begin + blockIdx.x * kBinWarps * 32 < e_end) sh_first_blk = atomicAdd(&a.totals->n_stage_blocks, (uint32_t)kBinWarps);
  __syncthreads();
  uint32_t blk = 0, blk_used = kStageBlock;  // current staging block of this warp (none yet)
  bool have_blk = false, stage_full = false;
  for (uint32_t base = e_begin + (blockIdx.x * kBinWarps + w) * 32; base < e_end; base += stride) {
    const uint32_t e = base + lane;
    int nb = 0;
    if (e < e_end) {
      const uint32_t pid = a.edge_pid[e];
      {
        const int4 ed = a.edges[e];
        const PathRec rec = a.path_rec[pid];
        const int bw = rec.wh & 0xffff, bh = rec.wh >> 16, by0 = rec.xy0 >> 16;
        if (bw) {
          const int ylo = min(ed.y, ed.w), yhi = max(ed.y, ed.w);
          int b_first = ylo >> 12, b_last = (ed.y == ed.w) ? b_first : (yhi - 1) >> 12;
          // rows outside the path's grid are outside the viewport (the bbox covers every edge of the path)
          b_first = max(b_first, by0);
          b_last = min(b_last, by0 + bh - 1);
          if (b_first <= b_last) {
            nb = b_last - b_first + 1;
            sh_edge[w][lane] = ed;
            sh_path[w][lane] = make_uint4(rec.xy0, (uint32_t)bw | ((uint32_t)b_first << 16), a.path_slot_off[pid], pid);
            sh_cover[w][lane] = (rec.info >> 16) * tiles;
          }
        }
      }
    }
    int incl = nb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - nb;
    __syncwarp();
    for (int k0 = 0; k0 < total; k0 += 32) {
      // ---- level 1: lane <-> (edge, band) ----
      const int k = min(k0 + (int)lane, total - 1);
      int o = 0;  // owner = the last lane whose first item is <= k (lanes without items share their successor's start)
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        int v = __shfl_sync(0xffffffffu, excl, o + step);
        if (v <= k) o += step;
      }
      const int first = __shfl_sync(0xffffffffu, excl, o);
      uint32_t nc = 0;
      if (k0 + (int)lane < total) {
        const int4 ed = sh_edge[w][o];
        const uint4 pp = sh_path[w][o];
        const int bx0 = pp.x & 0xffff, by0 = pp.x >> 16, bw = pp.y & 0xffff;
        const bool small = max(abs(ed.z - ed.x), abs(ed.w - ed.y)) <= kMaxLenFx;
        const int b = (int)(pp.y >> 16) + (k - first);
        const BandPiece piece = small ? band_setup<true>(a, ed.x, ed.y, ed.z, ed.w, bx0, by0, bw, pp.z, b)
                                      : band_setup<false>(a, ed.x, ed.y, ed.z, ed.w, bx0, by0, bw, pp.z, b);
        if (piece.c1 >= piece.c0) {
          nc = (uint32_t)(piece.c1 - piece.c0 + 1);
          sh_piece[w][lane] = make_int4(piece.xs, piece.ys, piece.xe, piece.ye);
          // .z = index of the band's first tile in the cover map (the band top follows from the piece: see level 2)
          sh_pmeta[w][lane] = make_uint4((uint32_t)piece.c0, piece.row_base + (uint32_t)(piece.c0 - bx0),
                                         sh_cover[w][o] + (uint32_t)b * (uint32_t)a.tiles_x, pp.w | (small ? 0x80000000u : 0u));
        }
      }
      uint32_t cincl = nc;
#pragma unroll
      for (int o2 = 1; o2 < 32; o2 <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, cincl, o2);
        if ((int)lane >= o2) cincl += t;
      }
      const uint32_t ctotal = __shfl_sync(0xffffffffu, cincl, 31);
      if (ctotal == 0) continue;
      const uint32_t cexcl = cincl - nc;
      __syncwarp();
      // ---- level 2: lane <-> (piece, column) ----
      for (uint32_t i0 = 0; i0 < ctotal; i0 += 32) {
        const uint32_t idx = min(i0 + lane, ctotal - 1);
        int o2 = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          uint32_t v = __shfl_sync(0xffffffffu, cexcl, o2 + step);
          if (v <= idx) o2 += step;
        }
        const uint32_t cfirst = __shfl_sync(0xffffffffu, cexcl, o2);
        bool keep = false;
        unsigned long long rc = 0;
        uint32_t slot = 0, pid = 0;
        if (i0 + lane < ctotal) {
          const uint4 pm = sh_pmeta[w][o2];
          const uint32_t j = idx - cfirst;
          pid = pm.w & 0x7fffffffu;
          slot = pm.y + j;
          // occlusion culling: an opaque path of a chunk above covers this tile completely
          if (__ldg(a.tile_cover + pm.z + pm.x + j) <= pid) {
            const int4 pc = sh_piece[w][o2];
            // a band piece lies in [Yt, Yt + 4096) with at most its lower end on the bottom line
            const int Yt = min(pc.y, pc.w) & ~(kTileFx - 1);
            keep = (pm.w >> 31) ? column_record<true>(pc.x, pc.y, pc.z, pc.w, Yt, (int)(pm.x + j), rc)
                                : column_record<false>(pc.x, pc.y, pc.z, pc.w, Yt, (int)(pm.x + j), rc);
          }
        }
        const uint32_t kmask = __ballot_sync(0xffffffffu, keep);
        if (kmask == 0) continue;
        const uint32_t kcount = __popc(kmask);
        if (!have_blk || blk_used + kcount > kStageBlock) {  // next staging block of this warp
          if (have_blk && !stage_full && lane == 0) a.stage_used[blk] = blk_used;
          if (!have_blk) {
            blk = sh_first_blk + w;  // the first block of every warp came from one atomic per thread block
          } else {
            uint32_t nb2 = 0;
            if (lane == 0) nb2 = atomicAdd(&a.totals->n_stage_blocks, 1u);
            blk = __shfl_sync(0xffffffffu, nb2, 0);
          }
          if (blk >= cap_blocks) {
            if (lane == 0) atomicOr(&a.totals->overflow_stage, 1u);
            stage_full = true;
          }
          blk_used = 0;
          have_blk = true;
        }
        if (keep) {
          atomicAdd(&a.slot_count[slot], 1u);
          if (!stage_full) {
            // streaming store: staging is written once here and read once by k_scatter
            const uint32_t pos = blk * kStageBlock + blk_used + __popc(kmask & ((1u << lane) - 1u));
            __stcs(a.stage + pos, make_uint4((uint32_t)rc, (uint32_t)(rc >> 32), slot, pid));
          }
        }
        blk_used += kcount;
      }
      __syncwarp();
    }
    __syncwarp();
  }
  if (have_blk && !stage_full && lane == 0) a.stage_used[blk] = blk_used;
}

// K2, pass 2 of 2: staged records -> their slots.  slot_off holds the END of the slot's record range (relative to the
// path's base); the counts of pass 1 double as cursors and run back down to zero (order inside a slot is irrelevant:
// coverage accumulation is integer).
__global__ void k_scatter(RenderArgs a) {
  pdl_enter();
  if (a.totals->overflow | a.totals->overflow_stage) return;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t nblk = min(a.totals->n_stage_blocks, a.caps.stage / kStageBlock);
  for (uint32_t b = warp; b < nblk; b += nwarps) {
    const uint32_t used = a.stage_used[b];
    const uint4 *st = a.stage + (size_t)b * kStageBlock;
    for (uint32_t i = lane; i < used; i += 32) {
      const uint4 e = __ldcs(st + i);
      const uint32_t pos = a.path_rec_base[e.w] + a.slot_off[e.z] - atomicSub(&a.slot_count[e.z], 1u);
      a.records[pos] = (unsigned long long)e.x | ((unsigned long long)e.y << 32);
    }
  }
}

// Per path instance, over the slots of its tile grid (row-major):
//   * prefix sum of the backdrop deltas along each tile row (winding number at the left edge of every tile);
//   * inclusive prefix sum of the record counts -> slot_off (END of each slot's record range, relative to the path)
//     and the path's total -> path_rec_base, which a scan over the PATHS then turns into the path's first record:
//     record space is allocated path by path, so there is no global scan over the (much longer) slot array.
// Grids of up to kBackdropSmall slots: one warp per path, 32 slots per step, coalesced.  Larger grids: one block
// per path (second kernel), so that a full-screen path is not a serial tail.
// First record of a path instance: record space is handed out path by path from one cursor (the placement of the
// paths in the record buffer varies from run to run; nothing observable depends on it).
__device__ __forceinline__ uint32_t alloc_records(const RenderArgs &a, uint32_t total) {
  if (total == 0) return 0u;
  const uint32_t base = atomicAdd(&a.totals->n_records, total);
  if (base + total > a.caps.records || base + total < base) atomicOr(&a.totals->overflow, 4u);
  return base;
}

// One warp, one path instance of up to kBackdropSmall slots, 32 slots per step (4 steps' loads in flight).
//   BACKDROP: prefix sum of the backdrop deltas along each tile row, and - for opaque paths - an atomicMax on
//             tile_cover for every slot without records and with a non-zero winding number (a full-tile cover);
//   COUNT:    inclusive prefix of the record counts -> slot_off, total -> record allocation.
template <bool BACKDROP, bool COUNT>
__device__ __forceinline__ void small_path_scan(const RenderArgs &a, uint32_t pid, const uint4 rec, uint32_t lane, uint32_t *cover,
                                                uint32_t *cover_bits) {
  const int bx0 = rec.x & 0xffff, by0 = rec.x >> 16, bw = rec.y & 0xffff, bh = rec.y >> 16;
  const bool opaque = (rec.z >> 8) & 1u;
  const int n = bw * bh;
  const uint32_t s0 = a.path_slot_off[pid];
  int32_t *bd = a.slot_backdrop + s0;  // winding deltas posted by k_bin: read, then zeroed again (self-cleaning)
  int32_t *wind = a.slot_wind + s0;    // winding number at the left edge of every tile (what k_fine reads)
  const uint32_t *cnt = a.slot_count + s0;
  uint32_t *end = a.slot_off + s0;
  int carry = 0, carry_row = -1;
  uint32_t ccarry = 0;
  for (int i0 = 0; i0 < n; i0 += 128) {
    uint32_t c4[4];
    int v4[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      int i = i0 + 32 * u + (int)lane;
      c4[u] = i < n ? cnt[i] : 0u;
      v4[u] = (BACKDROP && i < n) ? bd[i] : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      int i = i0 + 32 * u + (int)lane;
      if (i0 + 32 * u >= n) break;
      bool ok = i < n;
      if (COUNT) {
        uint32_t cc = c4[u];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          uint32_t t = __shfl_up_sync(0xffffffffu, cc, o);
          if ((int)lane >= o) cc += t;
        }
        cc += ccarry;
        if (ok) end[i] = cc;
        ccarry = __shfl_sync(0xffffffffu, cc, 31);
      }
      if (BACKDROP) {
        int row = ok ? i / bw : -2;
        int col = ok ? i - row * bw : 0;
        int v = v4[u];
        if (bw > 1) {
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, v, o);
            if ((int)lane >= o && col >= o) v += t;
          }
          if (row == carry_row) v += carry;
          carry = __shfl_sync(0xffffffffu, v, 31);
          carry_row = __shfl_sync(0xffffffffu, row, 31);
        }
        if (ok) {
          wind[i] = v;
          if (v4[u] != 0) bd[i] = 0;
        }
        // an opaque path covers this tile completely: it hides the chunks below
        if (ok && opaque && c4[u] == 0 && v != 0) {
          atomicMax(cover + (by0 + row) * a.tiles_x + bx0 + col, pid + 1u);
          atomicOr(cover_bits + (size_t)(by0 + row) * a.cover_words + ((bx0 + col) >> 5), 1u << ((bx0 + col) & 31));
        }
      }
    }
  }
  if (COUNT && lane == 0) a.path_rec_base[pid] = alloc_records(a, ccarry);
}

// The same for a path instance with a larger tile grid, by one block: warps over rows for the backdrop, a block-wide
// prefix for the counts.
template <bool BACKDROP, bool COUNT>
__device__ __forceinline__ void big_path_scan(const RenderArgs &a, uint32_t pid, const uint4 rec, uint32_t *cover, uint32_t *cover_bits,
                                              uint32_t *sh) {
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int bx0 = rec.x & 0xffff, by0 = rec.x >> 16, bw = rec.y & 0xffff, bh = rec.y >> 16;
  const bool opaque = (rec.z >> 8) & 1u;
  const uint32_t s0 = a.path_slot_off[pid];
  if (BACKDROP) {
    for (int row = (int)w; row < bh; row += (int)nw) {
      int32_t *q = a.slot_backdrop + s0 + row * bw;
      const uint32_t *cq = a.slot_count + s0 + row * bw;
      int carry = 0;
      for (int x0 = 0; x0 < bw; x0 += 32) {
        int x = x0 + (int)lane;
        const int delta = x < bw ? q[x] : 0;
        int v = delta;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, v, o);
          if ((int)lane >= o) v += t;
        }
        v += carry;
        if (x < bw) {
          a.slot_wind[s0 + row * bw + x] = v;
          if (delta != 0) q[x] = 0;  // self-cleaning
          if (opaque && v != 0 && cq[x] == 0) {
            atomicMax(cover + (by0 + row) * a.tiles_x + bx0 + x, pid + 1u);
            atomicOr(cover_bits + (size_t)(by0 + row) * a.cover_words + ((bx0 + x) >> 5), 1u << ((bx0 + x) & 31));
          }
        }
        carry = __shfl_sync(0xffffffffu, v, 31);
      }
    }
  }
  if (COUNT) {
    const uint32_t n = (uint32_t)(bw * bh);
    uint32_t carry = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += blockDim.x * 4) {
      const uint32_t i = i0 + threadIdx.x * 4;
      uint32_t cc[4];
#pragma unroll
      for (int k = 0; k < 4; k++) cc[k] = i + k < n ? a.slot_count[s0 + i + k] : 0u;
      uint32_t tile_total;
      uint32_t run = carry + block_exclusive(cc[0] + cc[1] + cc[2] + cc[3], sh, &tile_total);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        run += cc[k];
        if (i + k < n) a.slot_off[s0 + i + k] = run;
      }
      carry += tile_total;
    }
    if (threadIdx.x == 0) a.path_rec_base[pid] = alloc_records(a, carry);
    __syncthreads();
  }
}

constexpr int kCoverBigBlocks = kNumSM * 8;  // blocks that serve the large-grid paths

// Per depth chunk, after its binning (a path is binned in its own chunk only, so its counts are final): winding
// numbers (backdrop prefix) of the chunk's visible paths, the tiles they cover opaquely, and the inclusive prefix of
// their record counts -> slot_off (END of each slot's record range, relative to the path) + the path's first record
// (alloc_records): record space is allocated path by path, so there is no global scan over the (much longer) slot
// array.  grid = small-path blocks + kCoverBigBlocks.
__global__ void __launch_bounds__(256) k_cover(RenderArgs a, uint32_t c) {
  pdl_enter();
  if (a.totals->overflow) return;
  __shared__ uint32_t sh[32];
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t tiles = (uint32_t)(a.tiles_x * a.tiles_y);
  const uint32_t small_blocks = gridDim.x - kCoverBigBlocks;
  if (blockIdx.x < small_blocks) {
    const uint32_t n_small = a.totals->n_small_chunk[c];  // the chunk's visible paths with small grids (k_path_alive)
    const uint32_t nwarps = (small_blocks * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_small; i += nwarps) {
      const uint32_t pid = a.small_chunk[i];
      const uint4 rec = __ldg(reinterpret_cast<const uint4 *>(a.path_rec + pid));
      small_path_scan<true, true>(a, pid, rec, lane, a.tile_cover + (rec.z >> 16) * tiles,
                                  a.cover_bits + (size_t)(rec.z >> 16) * a.tiles_y * a.cover_words);
    }
  } else {
    const uint32_t n_big = a.totals->n_big_chunk[c];  // ... with large grids
    for (uint32_t bi = blockIdx.x - small_blocks; bi < n_big; bi += kCoverBigBlocks) {
      const uint32_t pid = a.big_chunk[bi];
      const uint4 rec = __ldg(reinterpret_cast<const uint4 *>(a.path_rec + pid));
      big_path_scan<true, true>(a, pid, rec, a.tile_cover + (rec.z >> 16) * tiles,
                                a.cover_bits + (size_t)(rec.z >> 16) * a.tiles_y * a.cover_words, sh);
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

// over_masked(dst, src, m) of the oracle - s = (m == 255) ? src : MUL_UN8x4(src, m); s.a == 255 -> s; s == 0 -> dst;
// else s + MUL_UN8x4(dst, 255 - s.a) - is computed in k_fine without the case analysis: the shortcuts are identities of
// the general form (MUL_UN8(x, 255) = x, MUL_UN8(x, 0) = 0), so s + MUL_UN8x4(dst, 255 - s.a) with s = MUL_UN8x4(src, m)
// is the same value in every case.

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
  float xt = fmaf(ytf - yaf, slope, xaf), xm = fmaf(ybmf - yaf, slope, xaf);
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
    float a1 = fmaf(u1 * u1, inv2w, fmaxf(fi1 - xmax, 0.0f));
    float f = a1 - a0;
    f = fminf(fmaxf(f, 0.0f), 1.0f);
    int c = __float2int_rn(Df * f);
    atomicAdd(&acc_row[i], c - prev);
    prev = c;
  }
  if (iend < 16) atomicAdd(&acc_row[iend], D - prev);
}

// a mod n in [0, n) for n > 0.  The quotient is estimated in float (inv_n = 1 / n from path setup) and corrected
// against the exact integer remainder: for |a| < 2^22 the estimate is off by at most one, so the result is the
// mathematical floor-mod (the oracle's floormod) without the integer division subroutine.
__device__ __forceinline__ int floormod(int a, int n, float inv_n) {
  if (a >= 0 && a < n) return a;
  if (abs(a) < (1 << 22)) {
    const int q = __float2int_rd((float)a * inv_n);
    int r = a - q * n;
    if (r < 0) r += n;
    if (r >= n) r -= n;
    return r;
  }
  const int r = a % n;
  return r < 0 ? r + n : r;
}

// Paint of device pixel (X, Y).  Every a*b+c is an explicit fused multiply-add, in the same places as the oracle's
// eval_paint (C fmaf is correctly rounded, fmaf() here compiles to FFMA; the rest of this file is compiled without
// contraction), so the colours are the oracle's bit for bit.
__device__ uint32_t eval_paint(uint32_t type, const PaintInst &p, int X, int Y) {
  const float xc = (float)X + 0.5f, yc = (float)Y + 0.5f;
  const float gx = fmaf(p.inv[0], xc, fmaf(p.inv[2], yc, p.inv[4]));
  const float gy = fmaf(p.inv[1], xc, fmaf(p.inv[3], yc, p.inv[5]));
  if (type == PAINT_BITMAP) {
    // box(rx) x box(ry) footprint around the sample point, texel by texel, rows outer / columns inner (the oracle's
    // summation order).  (Textures of four floats per texel - no byte -> float conversion per tap - were measured: 3.5 %
    // fewer instructions, no change on the 10 k shapes stream, and 23 % slower on the minified bitmaps of config 4,
    // whose 9 x 9 footprints then move four times the bytes through the texture cache.)  Wrapped texel indices are carried along instead of taking a modulo per tap; 1 / rx and
    // 1 / ry come from path setup (p.focal / p.omf are reused for them: same IEEE division, done once per path).
    cudaTextureObject_t tex = (cudaTextureObject_t)p.ptr;
    const float hrx = p.rx * 0.5f, hry = p.ry * 0.5f;
    const float irx = p.focal, iry = p.omf;
    const float lox = gx - hrx, hix = gx + hrx, loy = gy - hry, hiy = gy + hry;
    const int i0 = (int)floorf(lox), j0 = (int)floorf(loy);
    const bool rep = p.repeating != 0;
    const int ii0 = rep ? floormod(i0, p.bw, p.inv_bw) : i0;
    int jj = rep ? floormod(j0, p.bh, p.inv_bh) : j0;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const int ncol = (int)ceilf(hix) - i0;  // texel columns i0 .. i0 + ncol - 1 are those with (float)i < hix
    const int nrow = (int)ceilf(hiy) - j0;  // texel rows likewise
    if (ncol <= 2 && nrow <= 2) {
      // bilinear footprint (no minification: a 1 x 1 texel box): at most 2 x 2 taps, straight-line code.  Same taps, same
      // weights, same order (rows outer, columns inner) as the general loops below.
      float wx[2], wy[2], xs[2], ys[2];
      bool vx[2], vy[2];
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int i = i0 + c;
        int t = ii0 + c;
        if (rep && t >= p.bw) t -= p.bw;
        xs[c] = (float)t + 0.5f;
        vx[c] = c < ncol && (rep || (i >= 0 && i < p.bw));
        wx[c] = fmaxf(fminf(hix, (float)(i + 1)) - fmaxf(lox, (float)i), 0.0f) * irx;
        const int j = j0 + c;
        int u = jj + c;
        if (rep && u >= p.bh) u -= p.bh;
        ys[c] = (float)u + 0.5f;
        vy[c] = c < nrow && (rep || (j >= 0 && j < p.bh));
        wy[c] = fmaxf(fminf(hiy, (float)(j + 1)) - fmaxf(loy, (float)j), 0.0f) * iry;
      }
#pragma unroll
      for (int rr = 0; rr < 2; rr++) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
          if (!(vy[rr] && vx[c])) continue;
          const uchar4 t = tex2D<uchar4>(tex, xs[c], ys[rr]);
          const float wgt = wx[c] * wy[rr];
          acc0 = fmaf(wgt, (float)t.x, acc0);
          acc1 = fmaf(wgt, (float)t.y, acc1);
          acc2 = fmaf(wgt, (float)t.z, acc2);
          acc3 = fmaf(wgt, (float)t.w, acc3);
        }
      }
    } else if (ncol <= 3) {
      // the usual footprints (bilinear, 2x minification): column weights, wrapped column indices and validity are
      // computed once per pixel instead of once per tap; the taps are visited in the same order (rows outer, columns
      // inner) with the same operands, so the sums are the oracle's
      float wxc[3];
      float xcs[3];
      bool vc[3];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const int i = i0 + c;
        int t = ii0 + c;
        if (rep && t >= p.bw) t -= p.bw;  // bw >= 1 and c <= 2: at most two wraps
        if (rep && t >= p.bw) t -= p.bw;
        xcs[c] = (float)t + 0.5f;
        vc[c] = c < ncol && (rep || (i >= 0 && i < p.bw));
        const float vl = fmaxf(lox, (float)i), vh = fminf(hix, (float)(i + 1));
        wxc[c] = fmaxf(vh - vl, 0.0f) * irx;
      }
      for (int j = j0; (float)j < hiy; j++) {
        const int jcur = jj;
        jj++;
        if (rep && jj == p.bh) jj = 0;
        if (!rep && (j < 0 || j >= p.bh)) continue;
        const float wl = fmaxf(loy, (float)j), wh = fminf(hiy, (float)(j + 1));
        const float wy = fmaxf(wh - wl, 0.0f) * iry;
        const float ycs = (float)jcur + 0.5f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          if (!vc[c]) continue;
          const uchar4 t = tex2D<uchar4>(tex, xcs[c], ycs);
          const float wgt = wxc[c] * wy;
          acc0 = fmaf(wgt, (float)t.x, acc0);
          acc1 = fmaf(wgt, (float)t.y, acc1);
          acc2 = fmaf(wgt, (float)t.z, acc2);
          acc3 = fmaf(wgt, (float)t.w, acc3);
        }
      }
    } else {
      for (int j = j0; (float)j < hiy; j++) {
        const int jcur = jj;
        jj++;
        if (rep && jj == p.bh) jj = 0;
        if (!rep && (j < 0 || j >= p.bh)) continue;
        const float wl = fmaxf(loy, (float)j), wh = fminf(hiy, (float)(j + 1));
        const float wy = fmaxf(wh - wl, 0.0f) * iry;
        int ii = ii0;
        for (int i = i0; (float)i < hix; i++) {
          const int icur = ii;
          ii++;
          if (rep && ii == p.bw) ii = 0;
          if (!rep && (i < 0 || i >= p.bw)) continue;
          const float vl = fmaxf(lox, (float)i), vh = fminf(hix, (float)(i + 1));
          const float wx = fmaxf(vh - vl, 0.0f) * irx;
          const uchar4 t = tex2D<uchar4>(tex, (float)icur + 0.5f, (float)jcur + 0.5f);
          const float wgt = wx * wy;
          acc0 = fmaf(wgt, (float)t.x, acc0);
          acc1 = fmaf(wgt, (float)t.y, acc1);
          acc2 = fmaf(wgt, (float)t.z, acc2);
          acc3 = fmaf(wgt, (float)t.w, acc3);
        }
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
    t = fmaf(gx, 1.0f / 32768.0f, 0.5f);
  } else {
    const float nx = gx * (1.0f / 16384.0f), ny = gy * (1.0f / 16384.0f);
    const float dx = nx - p.focal;
    const float c = p.omf * (ny * ny);
    const float disc = fmaf(dx, dx, c);
    const float s = sqrtf(disc);
    const float num = fmaf(p.focal, dx, s);
    t = num * p.rx;  // gradients: rx carries 1 / omf
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
  // the ramp holds premultiplied RGBA8 at t = (k + 1/2) / kRampSize: one load, no interpolation
  const int i = min(max(__float2int_rd(t * (float)kRampSize), 0), kRampSize - 1);
  return __ldg(reinterpret_cast<const uint32_t *>(p.ptr) + i);
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
  if (!__ldg(a.path_alive + pid)) return pr;  // hidden everywhere: never binned, its slots hold nothing
  const uint32_t local = (uint32_t)(ly * bw + lx);
  const uint32_t slot = __ldg(a.path_slot_off + pid) + local;
  const uint32_t rec_base = __ldg(a.path_rec_base + pid);
  pr.o0 = rec_base + (local ? __ldg(a.slot_off + slot - 1) : 0u);
  pr.o1 = rec_base + __ldg(a.slot_off + slot);
  pr.bd = __ldg(a.slot_wind + slot);
  pr.info = rc.z;
  pr.color = rc.w;
  pr.hit = (pr.o1 > pr.o0) || (pr.bd != 0);
  return pr;
}

// 8-bit coverage masks m[0..7] of this lane's 8 pixels (row = lane & 15, columns (lane >> 4) * 8 ..) for one slot:
// stages the slot's records [o0, o1) into the warp's 16x16 Q16 signed-area accumulator (shared-memory integer
// atomics of prefix differences; crossings of the tile's left boundary go to a 17-entry column), then every lane
// prefix-sums its 8 accumulators, applies the backdrop and the non-zero rule min(|sum|, 1), and clears what it read.
__device__ __forceinline__ void slot_coverage(const RenderArgs &a, uint32_t o0, uint32_t o1, int bd, int *acc, int *cross,
                                              uint32_t lane, uint32_t m[8]) {
  const int row = lane & 15, half = lane >> 4;
  const uint32_t nrec = o1 - o0;
  bool any_cross = false;
  // Warp-cooperative expansion: 32 records per round; every record counts the pixel rows it touches, and the
  // (record, row) pairs of the round are spread evenly over the lanes (prefix sum + shuffle search for the owner), so a
  // short piece of a flattened curve does not leave the lanes of its fifteen other rows idle.
  for (uint32_t base = 0; base < nrec; base += 32) {
    unsigned long long rc = 0;
    int cnt = 0, r_first = 0;
    if (base + lane < nrec) {
      rc = __ldg(a.records + o0 + base + lane);
      const int ya = (int)((rc >> 13) & 0x1fff), yb = (int)((rc >> 39) & 0x1fff);
      const int fs = (int)((rc >> 52) & 1), fe = (int)((rc >> 53) & 1);
      if (fs | fe) {  // the record enters or leaves through the tile's left boundary: crossing term of that height
        const int yc = fs ? ya : yb, sgn = fs ? -1 : 1;
        const int r0 = min(yc >> 8, 15);
        const int hq = min(max((r0 + 1) * 256 - yc, 0), 256) * 256;
        atomicAdd(&cross[r0], sgn * hq);
        atomicAdd(&cross[r0 + 1], sgn * (65536 - hq));
        any_cross = true;
      }
      if (ya != yb) {
        const int ylo = min(ya, yb), yhi = max(ya, yb);
        r_first = ylo >> 8;
        cnt = ((yhi - 1) >> 8) - r_first + 1;
      }
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - cnt;
    for (int k0 = 0; k0 < total; k0 += 32) {
      const int k = min(k0 + (int)lane, total - 1);
      int o = 0;  // owner = the last lane whose first pair is <= k
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int v = __shfl_sync(0xffffffffu, excl, o + step);
        if (v <= k) o += step;
      }
      const int first = __shfl_sync(0xffffffffu, excl, o);
      const int r = __shfl_sync(0xffffffffu, r_first, o) + (k - first);
      const unsigned long long orc = __shfl_sync(0xffffffffu, rc, o);
      if (k0 + (int)lane < total)
        accumulate_row((int)(orc & 0x1fff), (int)((orc >> 13) & 0x1fff), (int)((orc >> 26) & 0x1fff), (int)((orc >> 39) & 0x1fff), r,
                       acc + r * kAccStride);
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
      if ((int)(lane & 15) >= o) cr += t;
    }
    basev += cr;
  }
  __syncwarp();
  *reinterpret_cast<int4 *>(acc + row * kAccStride + half * 8) = make_int4(0, 0, 0, 0);
  *reinterpret_cast<int4 *>(acc + row * kAccStride + half * 8 + 4) = make_int4(0, 0, 0, 0);
  if (any_cross && lane < 20) cross[lane] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int sm = v[i] + basev;
    sm = sm < 0 ? -sm : sm;
    sm = min(sm, 65536);
    m[i] = ((uint32_t)sm * 255u + 32768u) >> 16;
  }
  __syncwarp();
}

// Coverage of a path that may overlap itself (stroke outlines).  The area integral above, clamped per pixel, over-covers
// where two parts of one path overlap inside a partly covered pixel; here the non-zero rule is applied per
// sub-scanline, the way Cairo's scan converter does it (15 sub-rows per pixel row, exact horizontal extents in
// 1/256 px).  Same definition as the oracle's sampled_coverage: sample lines through the centres of the sub-rows,
// a record covers the lines in [y_lo, y_hi) of its own extent, crossings walked in ascending (x, record) order.
// One lane per sub-scanline (240 of them, 8 rounds); the slot's records are decoded once into shared memory.
constexpr int kSubRows = 15;
constexpr int kMaxSampled = 96;  // more records in one slot: the area integral is used (oracle: SWFO_MAX_SAMPLED)

__device__ __noinline__ uint2 slot_coverage_sampled(const unsigned long long *__restrict__ records, uint32_t nrec, int bd, int *acc,
                                                    int4 *stage, uint32_t lane) {
  for (uint32_t j = lane; j < nrec; j += 32) {
    const unsigned long long rc = __ldg(records + j);
    const int xa = (int)(rc & 0x1fff), ya = (int)((rc >> 13) & 0x1fff), xb = (int)((rc >> 26) & 0x1fff), yb = (int)((rc >> 39) & 0x1fff);
    const int fs = (int)((rc >> 52) & 1), fe = (int)((rc >> 53) & 1);
    const float slope = ya != yb ? (float)(xb - xa) / (float)((yb - ya) * kSubRows) : 0.0f;
    stage[j] = make_int4(xa | (xb << 16), (ya * kSubRows) | (fs << 16) | (fe << 17), yb * kSubRows, __float_as_int(slope));
  }
  __syncwarp();
  for (int k = (int)lane; k < 16 * kSubRows; k += 32) {
    const int Yk = k * 256 + 128;  // the sample line, in 1/(256 * 15) px
    int *arow = acc + (k / kSubRows) * kAccStride;
    int w = bd;  // winding at the tile's left edge on this line
    for (uint32_t i = 0; i < nrec; i++) {
      const int4 r = stage[i];
      const int fl = r.y >> 16;
      if (fl) {
        const int yc15 = (fl & 1) ? (r.y & 0xffff) : r.z;
        if (Yk >= yc15) w += (fl & 1) ? -1 : 1;
      }
    }
    if (w != 0) atomicAdd(&arow[0], 256);
    int last_x = -1, last_i = -1;
    for (;;) {
      int best_x = INT_MAX, best_i = -1;
      for (uint32_t i = 0; i < nrec; i++) {
        const int4 r = stage[i];
        const int ya15 = r.y & 0xffff, yb15 = r.z;
        if (Yk < min(ya15, yb15) || Yk >= max(ya15, yb15)) continue;
        const int xa = r.x & 0xffff, xb = r.x >> 16;
        float x = fmaf((float)(Yk - ya15), __int_as_float(r.w), (float)xa);
        x = fminf(fmaxf(x, (float)min(xa, xb)), (float)max(xa, xb));
        const int xi = __float2int_rn(x);
        if (!(xi > last_x || (xi == last_x && (int)i > last_i))) continue;
        if (xi < best_x) best_x = xi, best_i = (int)i;
      }
      if (best_i < 0) break;
      const int4 r = stage[best_i];
      const int w2 = w + (r.z > (r.y & 0xffff) ? 1 : -1);
      const int e = (w == 0 && w2 != 0) ? 1 : ((w != 0 && w2 == 0) ? -1 : 0);
      if (e) {
        const int p = best_x >> 8, f = best_x & 255;
        atomicAdd(&arow[p], e * (256 - f));
        atomicAdd(&arow[p + 1], e * f);
      }
      w = w2;
      last_x = best_x, last_i = best_i;
    }
  }
  __syncwarp();
  const int row = lane & 15, half = lane >> 4;
  int v[8];
  const int4 q0 = *reinterpret_cast<int4 *>(acc + row * kAccStride + half * 8);
  const int4 q1 = *reinterpret_cast<int4 *>(acc + row * kAccStride + half * 8 + 4);
  v[0] = q0.x, v[1] = q0.y, v[2] = q0.z, v[3] = q0.w, v[4] = q1.x, v[5] = q1.y, v[6] = q1.z, v[7] = q1.w;
#pragma unroll
  for (int i = 1; i < 8; i++) v[i] += v[i - 1];
  const int left = __shfl_sync(0xffffffffu, v[7], lane & 15);
  const int basev = half ? left : 0;
  __syncwarp();
  *reinterpret_cast<int4 *>(acc + row * kAccStride + half * 8) = make_int4(0, 0, 0, 0);
  *reinterpret_cast<int4 *>(acc + row * kAccStride + half * 8 + 4) = make_int4(0, 0, 0, 0);
  if (half) *reinterpret_cast<int4 *>(acc + row * kAccStride + 16) = make_int4(0, 0, 0, 0);  // crossings on the right boundary
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int c = min(max(v[i] + basev, 0), 256 * kSubRows);
    const uint32_t mi = (uint32_t)(34 * c + 256) >> 9;  // tor's GRID_AREA_TO_ALPHA for a 2 x 256 x 15 grid
    if (i < 4)
      lo |= mi << (8 * i);
    else
      hi |= mi << (8 * (i - 4));
  }
  __syncwarp();
  return make_uint2(lo, hi);  // eight 8-bit masks of this lane's pixels
}

// Four resident blocks per SM (64 registers per thread): measured best - 3 blocks at 80 registers 0.75 ms per launch,
// 4 at 64 0.54, 5 at 48 and 6 at 40 0.56 (1080p / 10 k shapes, 16 frames per launch).
// SAMPLED: the pass draws stroke outlines (RenderArgs.has_sampled) or carries colour transforms (RenderArgs.has_cx): the
// rare work.  The variant without them does not carry the call to slot_coverage_sampled (measured: its stack frame and
// register pressure cost the hot path 5 %) nor the per-pixel colour transform of gradients and bitmaps.
template <bool SAMPLED>
__global__ void __launch_bounds__(kFineWarps * 32, 4) k_fine(RenderArgs a, uint32_t slice, uint32_t frame_begin, uint32_t frame_end) {
  pdl_wait();  // (no launch_dependents: the successor's blocks would only squat in the slots the tail frees)
  if (a.totals->overflow | a.totals->overflow_stage) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.arena_dirty = 1u;  // see k_init
    return;
  }
  __shared__ int acc_sh[kFineWarps][16 * kAccStride];
  __shared__ int cross_sh[kFineWarps][20];
  __shared__ uint4 paint_sh[kFineWarps][sizeof(PaintInst) / 16];  // the paint instance the warp is compositing
  __shared__ int4 samp_sh[SAMPLED ? kFineWarps : 1][SAMPLED ? kMaxSampled : 1];  // decoded records of a slot of a stroke outline
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t n_hits = 0, n_recs = 0;  // statistics (swfr_stats.fine_*): per warp, posted once at the end
  int *acc = acc_sh[warp];
  int *cross = cross_sh[warp];
  for (int i = lane; i < 16 * kAccStride; i += 32) acc[i] = 0;
  if (lane < 20) cross[lane] = 0;
  __syncwarp();
  const int row = lane & 15, half = lane >> 4;
  const uint32_t tiles = (uint32_t)(a.tiles_x * a.tiles_y);
  const uint32_t total = tiles * (frame_end - frame_begin);  // this launch composites frames [frame_begin, frame_end)
  // (a warp leaves through the break below; the bound only states that no warp can take more than `total` tiles.  With
  // a counted loop ptxas fits the kernel in 64 registers without a single spill; with `while (true)` it spills 92 bytes
  // into the tile loop and a launch takes 0.61 instead of 0.54 ms)
  for (uint32_t taken = 0; taken < total; taken++) {
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(&a.totals->work[slice], 1u);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= total) break;
    uint32_t frame = frame_begin + w / tiles, tile = w % tiles;
    int ty = (int)(tile / (uint32_t)a.tiles_x), tx = (int)(tile - (uint32_t)ty * a.tiles_x);
    const int X0 = tx * kTile + half * 8, Y = ty * kTile + row;
    // candidates: the paint-ordered list of this tile's (row, column group)
    const uint32_t li = (frame * (uint32_t)a.tiles_y + (uint32_t)ty) * a.groups_x + (uint32_t)tx / kGroupTiles;
    const uint32_t p_begin = __ldg(a.list_off + li), p_end = __ldg(a.list_off + li + 1);

    // ---- pass 1: the last opaque full-tile cover hides everything painted before it (exact occlusion culling).
    // (Culling per pixel - also using partial opaque covers - was measured: on the 10 k shapes stream it removes 25 %
    // of the paint evaluations but costs more than that in extra coverage passes; profiles/, round 1.)
    uint32_t start = p_begin;
    for (uint32_t hi = p_end; hi > p_begin;) {
      uint32_t lo = hi - p_begin >= 32 ? hi - 32 : p_begin;
      uint32_t idx = lo + lane;
      uint32_t pid = idx < hi ? __ldg(a.list_items + idx) : 0u;
      Probe pr = probe_slot(a, pid, idx < hi, tx, ty);
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
    const uint32_t bg = __ldg(a.frame_bg + frame);
#pragma unroll
    for (int i = 0; i < 8; i++) px[i] = bg;
    for (uint32_t base = start; base < p_end; base += 32) {
      uint32_t idx = base + lane;
      uint32_t pid = idx < p_end ? __ldg(a.list_items + idx) : 0u;
      Probe pr = probe_slot(a, pid, idx < p_end, tx, ty);
      uint32_t mask = __ballot_sync(0xffffffffu, pr.hit);
      while (mask) {
        int src_lane = __ffs(mask) - 1;
        mask &= mask - 1;
        uint32_t o0 = __shfl_sync(0xffffffffu, pr.o0, src_lane);
        uint32_t o1 = __shfl_sync(0xffffffffu, pr.o1, src_lane);
        int bd = __shfl_sync(0xffffffffu, pr.bd, src_lane);
        uint32_t info = __shfl_sync(0xffffffffu, pr.info, src_lane);
        uint32_t color = __shfl_sync(0xffffffffu, pr.color, src_lane);
        uint32_t cur_pid = __shfl_sync(0xffffffffu, pid, src_lane);
        uint32_t type = info & 0xff;
        n_hits++;
        n_recs += o1 - o0;
        uint32_t m[8];
        if (o1 == o0) {
#pragma unroll
          for (int i = 0; i < 8; i++) m[i] = 255u;
        } else if (SAMPLED && ((info >> 9) & 1u) && o1 - o0 <= (uint32_t)kMaxSampled) {
          // a stroke outline: non-zero rule per sub-scanline (masks come back packed, four per word)
          const uint2 mp = slot_coverage_sampled(a.records + o0, o1 - o0, bd, acc, samp_sh[SAMPLED ? warp : 0], lane);
#pragma unroll
          for (int i = 0; i < 8; i++) m[i] = ((i < 4 ? mp.x : mp.y) >> (8 * (i & 3))) & 255u;
        } else {
          slot_coverage(a, o0, o1, bd, acc, cross, lane, m);
        }
        // ---- paint + blend ----
        if (type == PAINT_SOLID && o1 == o0) {
          // the whole tile inside a solid path: over_masked(px, color, 255) without its per-pixel case analysis
          const uint32_t ca = color >> 24;
          if (ca == 255u) {
#pragma unroll
            for (int i = 0; i < 8; i++) px[i] = color;
          } else {
#pragma unroll
            for (int i = 0; i < 8; i++) px[i] = color + mul_un8x4(px[i], 255u - ca);
          }
        } else if (type == PAINT_SOLID) {
          // partly covered tile of a solid path: over_masked without its case analysis.  The three shortcuts of
          // over_masked are identities of the general form (MUL_UN8(x, 255) = x, MUL_UN8(x, 0) = 0), and in an edge tile
          // some lane takes the general form at nearly every pixel anyway, so the branch-free form issues fewer
          // instructions than the divergent one (18 of 32 lanes were active in the blend arithmetic).
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const uint32_t sm = mul_un8x4(color, m[i]);
            px[i] = sm + mul_un8x4(px[i], 255u - (sm >> 24));
          }
        } else {
          // One copy of the paint code, eight trips.  px[] and the masks are only ever indexed statically (the pixels
          // rotate through px[0], the masks are packed into two words), so both stay in registers: a dynamic index
          // would put them in local memory for the whole kernel.
          // (staged in shared memory: one 16-byte load by each of five lanes instead of sixteen global loads per pixel)
          if (lane < (int)(sizeof(PaintInst) / 16))
            paint_sh[warp][lane] = __ldg(reinterpret_cast<const uint4 *>(a.paint_inst + cur_pid) + lane);
          __syncwarp();
          const PaintInst &pi = *reinterpret_cast<const PaintInst *>(paint_sh[warp]);
          uint32_t mlo = m[0] | (m[1] << 8) | (m[2] << 16) | (m[3] << 24);
          uint32_t mhi = m[4] | (m[5] << 8) | (m[6] << 16) | (m[7] << 24);
          // warp-uniform cases of the blend (the general form below is exact for all of them, see the solid paint):
          // every pixel of the tile fully covered, and an opaque paint on top of that
          // (kept in the spare bits of the paint type: bit 8 = fully covered, bit 9 = and opaque -> the source replaces)
          const uint32_t tf = type | (o1 == o0 ? (0x100u | ((info & 0x100u) << 1)) : 0u);
#pragma unroll 1
          for (int i = 0; i < 8; i++) {
            const uint32_t mm = mlo & 255u;
            uint32_t v = px[0];
            if (mm) {
              uint32_t src = eval_paint(tf & 0xffu, pi, X0 + i, Y);
              if (SAMPLED && pi.cx_on) src = cx_premul(src, pi.cx);
              if (tf & 0x200u) {
                v = src;
              } else {
                const uint32_t sm = (tf & 0x100u) ? src : mul_un8x4(src, mm);
                v = sm + mul_un8x4(v, 255u - (sm >> 24));
              }
            }
#pragma unroll
            for (int k = 0; k < 7; k++) px[k] = px[k + 1];
            px[7] = v;
            mlo = __funnelshift_r(mlo, mhi, 8);
            mhi >>= 8;
          }
          __syncwarp();  // the next paint instance overwrites paint_sh
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
  if (lane == 0 && n_hits) {
    atomicAdd(&a.totals->fine_hits, n_hits);
    atomicAdd(&a.totals->fine_records, n_recs);
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

// Straight RGBA8 rows (`stride` bytes apart) -> tight premultiplied 8-bit (what a Canvas stores); *translucent is set
// when any alpha is below 255.
__global__ void k_premultiply(const uint8_t *__restrict__ src, size_t stride, uint32_t w, uint32_t h, uint32_t *__restrict__ dst,
                              uint32_t *translucent) {
  const uint64_t n = (uint64_t)w * h, step = (uint64_t)gridDim.x * blockDim.x;
  bool any = false;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    const uint32_t y = (uint32_t)(i / w), x = (uint32_t)(i - (uint64_t)y * w);
    const uint8_t *s = src + (size_t)y * stride + 4 * (size_t)x;
    const uint32_t al = s[3];
    const uint32_t r = (s[0] * al + 127u) / 255u, g = (s[1] * al + 127u) / 255u, b = (s[2] * al + 127u) / 255u;
    dst[i] = r | (g << 8) | (b << 16) | (al << 24);
    any |= al != 255u;
  }
  if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) atomicOr(translucent, 1u);
}

// image/x-swf-bmp format 3 after inflate (decode-x-swf-bmp.ts:17-39): `colors` RGB triplets, then rows of 8-bit
// colour indices padded to a multiple of 4 bytes -> tight opaque RGBA8 (premultiplied == straight at alpha 255); an
// index past the table is opaque black (decode-x-swf-bmp.ts:35-36).
__global__ void k_xswfbmp_expand(const uint8_t *__restrict__ inflated, uint32_t colors, uint32_t w, uint32_t h, uint32_t padded,
                                 uint32_t *__restrict__ dst) {
  __shared__ uint32_t table[256];
  for (uint32_t c = threadIdx.x; c < 256; c += blockDim.x)
    table[c] = c < colors ? ((uint32_t)inflated[3 * c] | ((uint32_t)inflated[3 * c + 1] << 8) | ((uint32_t)inflated[3 * c + 2] << 16) |
                             0xff000000u)
                          : 0xff000000u;
  __syncthreads();
  const uint8_t *idx = inflated + 3 * (size_t)colors;
  const uint64_t n = (uint64_t)w * h, step = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    const uint32_t y = (uint32_t)(i / w), x = (uint32_t)(i - (uint64_t)y * w);
    dst[i] = table[idx[(size_t)y * padded + x]];
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
        uint32_t c = a.slot_off[s] - (s > base ? a.slot_off[s - 1] : 0u);
        if (c) atomicAdd(&counts[(by0 + y) * a.tiles_x + bx0 + x], c);
      }
  }
}

}  // namespace

// ======================================================================================================
// launchers
// ======================================================================================================

// Launch with the programmatic stream serialization attribute (see pdl_enter); SWFR_PDL=0 launches plainly.
static const bool g_pdl = [] {
  const char *e = getenv("SWFR_PDL");
  return !(e && atoi(e) == 0);
}();
template <class... P, class... A>
static void launch_k(void (*kern)(P...), dim3 grid, dim3 block, cudaStream_t st, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, P(args)...);
}

static void scan_u32(const uint32_t *src, uint32_t *dst, const uint32_t *n_ptr, uint32_t n_cap, uint32_t *tmp,
                     uint32_t *total_out, uint32_t total_cap, uint32_t *overflow, uint32_t bit, cudaStream_t st,
                     int &launches) {
  if (!n_ptr && n_cap <= kScanSmallMax) {
    launch_k(k_scan_small, dim3(1), dim3(1024), st, src, dst, n_cap, total_out, total_cap, overflow, bit);
    launches += 1;
    return;
  }
  launch_k(k_scan_partials, dim3(kScanBlocks), dim3(kScanThreads), st, src, n_ptr, n_cap, tmp);
  launch_k(k_scan_spine, dim3(1), dim3(1024), st, tmp, dst, n_ptr, n_cap, total_out, total_cap, overflow, bit);
  launch_k(k_scan_apply, dim3(kScanBlocks), dim3(kScanThreads), st, src, dst, n_ptr, n_cap, tmp);
  launches += 3;
}

const char *stage_name(int i) {
  static const char *names[kNumStages] = {"flatten_count", "scan_edges", "path_setup", "bin_chunks", "lists", "fine"};
  return (i >= 0 && i < kNumStages) ? names[i] : "?";
}

// ev (optional): kNumStages + 1 events recorded at the stage boundaries (profiling runs only).
int launch_render(const RenderArgs &a, cudaStream_t st, cudaEvent_t *ev, cudaEvent_t *slice_done) {
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
  launch_k(k_init, dim3(grid_for(a.n_paths)), dim3(T), st, a);
  launches++;
  // One depth chunk: every path is flattened, edges in segment order (the taps' mode).  More: only what is visible is
  // flattened, unordered, chunk by chunk.
  const bool ordered = a.n_chunks == 1;
  if (a.n_jobs) {  // outlines of morph-shape strokes, generated on the device for this pass' draws
    launch_k(k_stroke, dim3((a.n_jobs + 63) / 64), dim3(64), st, a);
    launches++;
  }
  if (a.n_seginst) {
    if (ordered) {
      launch_k(k_flatten_count, dim3(grid_for((uint64_t)a.n_items * 32)), dim3(T), st, a);
      launches++;
    }
  }
  mark(1);
  // edges: seg_edge_off (piece counts) -> exclusive offsets, total -> totals.n_edges
  if (ordered)
    scan_u32(a.seg_edge_off, a.seg_edge_off, nullptr, a.n_seginst, a.scan_tmp, &a.totals->n_edges, a.caps.edges, &a.totals->overflow,
             1u, st, launches);
  mark(2);
  if (a.n_paths) {
    launch_k(k_path_setup, dim3(grid_for(a.n_paths)), dim3(T), st, a);
    launches++;
  }
  scan_u32(a.path_slot_off, a.path_slot_off, nullptr, a.n_paths, a.scan_tmp, &a.totals->n_slots, a.caps.slots, &a.totals->overflow, 2u,
           st, launches);
  mark(3);
  // K1 emit + K2, depth chunk by depth chunk from the top one down: what a chunk covers opaquely hides the
  // geometry of the chunks below it (grid.y = frame; every kernel walks its frame's part of the chunk)
  if (a.n_seginst && a.n_paths) {
    const unsigned per_frame = std::max(1u, wide / std::max(1u, a.n_frames));
    const dim3 g2(per_frame, a.n_frames);
    for (uint32_t c = a.n_chunks; c-- > 0;) {
      launch_k(k_path_alive, dim3(g2), dim3(T), st, a, c);
      if (ordered) {
        launch_k(k_flatten_emit<true>, dim3(g2), dim3(kEmitWarps * 32), st, a, c);
        launch_k(k_bin<true>, dim3(g2), dim3(kBinWarps * 32), st, a, c);
      } else {
        launch_k(k_flatten_emit<false>, dim3(g2), dim3(kEmitWarps * 32), st, a, c);
        launch_k(k_bin<false>, dim3(wide), dim3(kBinWarps * 32), st, a, c);
      }
      launch_k(k_cover, dim3(wide + kCoverBigBlocks), dim3(T), st, a, c);
      launches += 4;
    }
    launch_k(k_scatter, dim3(wide), dim3(T), st, a);
    launches++;
  }
  mark(4);
  // candidate lists of the visible paths (for k_fine): row counts (from path setup) -> row lists -> group counts -> group lists
  scan_u32(a.row_count, a.row_off, nullptr, a.n_frames * (uint32_t)a.tiles_y, a.scan_tmp, &a.totals->n_rowent, a.caps.rows,
           &a.totals->overflow, 16u, st, launches);
  launch_k(k_alive_paths, dim3((unsigned)std::min<uint32_t>(a.n_frames, kNumSM * 8)), dim3(1024), st, a);
  launch_k(k_row_lists, dim3((unsigned)std::min<uint32_t>(a.n_frames * (uint32_t)a.tiles_y, kNumSM * 8)), dim3(kRowThreads), st, a);
  launches += 2;
  scan_u32(a.list_off, a.list_off, nullptr, a.n_lists, a.scan_tmp, &a.totals->n_list, a.caps.list, &a.totals->overflow, 8u, st,
           launches);
  launch_k(k_group_lists, dim3(grid_for((uint64_t)a.n_lists * 32)), dim3(T), st, a);
  launches++;
  mark(5);
  static const int fine_blocks = [] {
    const char *e = getenv("SWFR_FINE_BLOCKS_PER_SM");
    int v = e ? atoi(e) : 4;
    return v < 1 ? 1 : (v > 4 ? 4 : v);
  }();
  {
    const uint32_t fs = fine_slice_frames(a.n_frames), ns = fine_slices(a.n_frames);
    for (uint32_t k = 0; k < ns; k++) {
      if (a.has_sampled | a.has_cx)
        launch_k(k_fine<true>, dim3(kNumSM * fine_blocks), dim3(kFineWarps * 32), st, a, k, k * fs, std::min(a.n_frames, (k + 1) * fs));
      else
        launch_k(k_fine<false>, dim3(kNumSM * fine_blocks), dim3(kFineWarps * 32), st, a, k, k * fs, std::min(a.n_frames, (k + 1) * fs));
      launches++;
      if (slice_done) cudaEventRecord(slice_done[k], st);
    }
  }
  mark(6);
  return launches;
}

void launch_unpremultiply(const uint32_t *src, uint32_t *dst, uint64_t n_px, cudaStream_t st) {
  k_unpremultiply<<<kNumSM * 8, 256, 0, st>>>(src, dst, n_px);
}
void launch_premultiply(const uint8_t *src, size_t stride, uint32_t w, uint32_t h, uint32_t *dst, uint32_t *translucent,
                        cudaStream_t st) {
  k_premultiply<<<kNumSM * 8, 256, 0, st>>>(src, stride, w, h, dst, translucent);
}
void launch_xswfbmp_expand(const uint8_t *inflated, uint32_t colors, uint32_t w, uint32_t h, uint32_t padded, uint32_t *dst,
                           cudaStream_t st) {
  k_xswfbmp_expand<<<kNumSM * 8, 256, 0, st>>>(inflated, colors, w, h, padded, dst);
}
void launch_tile_counts(const RenderArgs &a, uint32_t frame, uint32_t *counts, cudaStream_t st) {
  k_tile_counts<<<kNumSM * 2, 256, 0, st>>>(a, frame, counts);
}

}  // namespace swfr
