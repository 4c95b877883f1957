// Stroke-to-fill expansion of morph-shape lines, written once for the host and the device.
//
// Reference semantics: ts/src/lib/renderers/canvas-renderer.ts:252-266 - the line path and its width are lerped at
// the draw's ratio, a zero width keeps the previous one, ctx.stroke() with round caps and round joins.  The outline of
// every sub-path is one closed contour (left side, end cap, right side, start cap) filled non-zero; curved pieces are
// offset as quadratics, arcs are made of <= 45 degree quadratics.  Only + - * / sqrt in double are used, in the order
// csrc/stroker.cpp (static shapes, at registration) and oracle/stroker.py use them, so the three produce the same
// float32 segments bit for bit: this file is that algorithm as a streaming generator - no containers, explicit stacks
// instead of recursion - so that one GPU thread can run it per morph draw (k_stroke) with no host geometry per draw.
// The host compiles the same code for the capacity estimate at registration and for the fallback when an outline
// outgrows its estimate.
#pragma once

#include <cmath>
#include <cstdint>

#include "host_types.h"

#if defined(__CUDACC__)
#define SWFR_HD __host__ __device__ __forceinline__
#else
#define SWFR_HD inline
#endif

namespace swfr {
namespace stroke {

constexpr double kCosSplit = 0.9659258262890683;  // cos 15 deg
constexpr double kCosArc = 0.7071067811865476;    // cos 45 deg

struct V2 {
  double x, y;
};
struct Span {  // an offsettable piece: line (curve = false) or quadratic with small turning
  V2 p0, c, p1, t0, t1;
  bool curve;
};

// One morph line as the device sees it.
struct LineDev {
  uint32_t cmd_first, cmd_count;  // into the morph command store
  double w0, w1;                  // width in twips, start / end state
  uint8_t color0[4], color1[4];
};
struct CmdDev {  // Command (host_types.h) without the vector around it
  int32_t type;  // LineTo = 0, CurveTo = 1, MoveTo = 2
  int32_t pad;
  double s[4], e[4];
};

SWFR_HD double lerp(double start, double end, double r) {  // canvas-renderer.ts:24-26
  double a = end * r;
  double b = 1.0 - r;
  double c = start * b;
  return a + c;
}

SWFR_HD bool unit(double dx, double dy, V2 &u) {
  double l = sqrt(dx * dx + dy * dy);
  if (l == 0.0) return false;
  u.x = dx / l;
  u.y = dy / l;
  return true;
}

// Receives the outline as float32 segments of one path; counts all of them, stores those that fit.
struct Sink {
  SegStatic *out;     // may be null (count only)
  uint32_t cap, n;    // room / segments produced so far (n may exceed cap: nothing is stored past it)
  uint32_t path;      // path index local to the draw
  float bounds[4];    // x_min, y_min, x_max, y_max over the control points (x_min > x_max: none)
  bool started;       // a contour is open
  float fx, fy, cx, cy;

  SWFR_HD void grow(float x, float y) {
    if (bounds[0] > bounds[2]) {
      bounds[0] = bounds[2] = x;
      bounds[1] = bounds[3] = y;
    } else {
      bounds[0] = fminf(bounds[0], x), bounds[2] = fmaxf(bounds[2], x);
      bounds[1] = fminf(bounds[1], y), bounds[3] = fmaxf(bounds[3], y);
    }
  }
  SWFR_HD void push(bool curve, float x0, float y0, float qx, float qy, float x1, float y1) {
    if (out && n < cap) {
      SegStatic g;
      g.p[0] = x0, g.p[1] = y0, g.p[2] = qx, g.p[3] = qy, g.p[4] = x1, g.p[5] = y1;
      g.path_flags = path | (curve ? 0x80000000u : 0u);
      out[n] = g;
    }
    n++;
    grow(x0, y0);
    grow(qx, qy);
    grow(x1, y1);
  }
  // the outline commands, in order: the first one of a contour is its starting point
  SWFR_HD void line_to(V2 p) {
    const float x = (float)p.x, y = (float)p.y;
    if (!started) {
      started = true;
      fx = cx = x, fy = cy = y;
      return;
    }
    if (x != cx || y != cy) push(false, cx, cy, cx, cy, x, y);
    cx = x, cy = y;
  }
  SWFR_HD void quad_to(V2 c, V2 p) {
    const float qx = (float)c.x, qy = (float)c.y, x = (float)p.x, y = (float)p.y;
    push(true, cx, cy, qx, qy, x, y);
    cx = x, cy = y;
  }
  SWFR_HD void close() {
    if (started && (fx != cx || fy != cy)) push(false, cx, cy, cx, cy, fx, fy);
    started = false;
  }
};

// Round join / cap piece from direction u to v around `center` (unit vectors, radius w): quadratics of <= 45 degrees.
SWFR_HD void arc(V2 center, V2 u, V2 v, double w, Sink &sink) {
  struct Node {
    V2 u, v;
    int depth;
  };
  Node st[8];
  int sp = 0;
  st[sp++] = Node{u, v, 0};
  while (sp > 0) {
    const Node nd = st[--sp];
    const double dot = nd.u.x * nd.v.x + nd.u.y * nd.v.y;
    if (dot < kCosArc && nd.depth < 6) {
      V2 mid;
      if (!unit(nd.u.x + nd.v.x, nd.u.y + nd.v.y, mid)) mid = V2{-nd.u.y, nd.u.x};
      st[sp++] = Node{mid, nd.v, nd.depth + 1};  // second half, visited after ...
      st[sp++] = Node{nd.u, mid, nd.depth + 1};  // ... the first half
      continue;
    }
    const double k = w / (1.0 + dot);
    sink.quad_to(V2{center.x + (nd.u.x + nd.v.x) * k, center.y + (nd.u.y + nd.v.y) * k},
                 V2{center.x + nd.v.x * w, center.y + nd.v.y * w});
  }
}

// Left-offset outline of one more span, with the join to the previous one ("left" is (ty, -tx) in y-down coordinates).
struct Side {
  bool first;
  V2 prev;
};
SWFR_HD void offset_span(const Span &s, double w, Side &side, Sink &sink) {
  const V2 n0{s.t0.y * w, -s.t0.x * w};
  const V2 n1{s.t1.y * w, -s.t1.x * w};
  const V2 start{s.p0.x + n0.x, s.p0.y + n0.y};
  if (side.first) {
    sink.line_to(start);
    side.first = false;
  } else {
    const V2 prev = side.prev;
    const double cross = prev.x * s.t0.y - prev.y * s.t0.x;
    const double dot = prev.x * s.t0.x + prev.y * s.t0.y;
    if (dot > 0.0 && fabs(cross) < 1e-12) {
      sink.line_to(start);
    } else if (cross > 0.0 || (cross == 0.0 && dot <= 0.0)) {  // outer side of the turn: round join
      arc(s.p0, V2{prev.y, -prev.x}, V2{s.t0.y, -s.t0.x}, w, sink);
    } else {  // inner side: through the vertex, keeps the contour's winding consistent
      sink.line_to(s.p0);
      sink.line_to(start);
    }
  }
  const V2 end{s.p1.x + n1.x, s.p1.y + n1.y};
  if (!s.curve) {
    sink.line_to(end);
  } else {
    const double dotn = s.t0.x * s.t1.x + s.t0.y * s.t1.y;
    const double k = w / (1.0 + dotn);
    sink.quad_to(V2{s.c.x + (s.t0.y + s.t1.y) * k, s.c.y + (-s.t0.x - s.t1.x) * k}, end);
  }
  side.prev = s.t1;
}

// The spans of one quadratic (p0, c, p1): halved until the tangent turns by at most 15 degrees (depth <= 8), visited
// first to last (REVERSE = false) or last to first with every span turned around (REVERSE = true).
template <bool REVERSE, class F>
SWFR_HD void quad_spans(V2 p0, V2 c, V2 p1, F &&f) {
  struct Node {
    V2 p0, c, p1;
    int depth;
  };
  Node st[10];
  int sp = 0;
  st[sp++] = Node{p0, c, p1, 0};
  while (sp > 0) {
    const Node nd = st[--sp];
    V2 t0{0, 0}, t1{0, 0};
    const bool h0 = unit(nd.c.x - nd.p0.x, nd.c.y - nd.p0.y, t0);
    const bool h1 = unit(nd.p1.x - nd.c.x, nd.p1.y - nd.c.y, t1);
    if (!h0 && !h1) {
      V2 ch;
      if (unit(nd.p1.x - nd.p0.x, nd.p1.y - nd.p0.y, ch)) {
        if (REVERSE)
          f(Span{nd.p1, nd.p0, nd.p0, V2{-ch.x, -ch.y}, V2{-ch.x, -ch.y}, false});
        else
          f(Span{nd.p0, nd.p0, nd.p1, ch, ch, false});
      }
      continue;
    }
    if (!h0) t0 = t1;
    if (!h1) t1 = t0;
    const double dot = t0.x * t1.x + t0.y * t1.y;
    if (dot >= kCosSplit || nd.depth >= 8) {
      if (REVERSE)
        f(Span{nd.p1, nd.c, nd.p0, V2{-t1.x, -t1.y}, V2{-t0.x, -t0.y}, true});
      else
        f(Span{nd.p0, nd.c, nd.p1, t0, t1, true});
      continue;
    }
    const V2 a{(nd.p0.x + nd.c.x) * 0.5, (nd.p0.y + nd.c.y) * 0.5};
    const V2 b{(nd.c.x + nd.p1.x) * 0.5, (nd.c.y + nd.p1.y) * 0.5};
    const V2 m{(a.x + b.x) * 0.5, (a.y + b.y) * 0.5};
    if (REVERSE) {
      st[sp++] = Node{nd.p0, a, m, nd.depth + 1};  // first half, visited after ...
      st[sp++] = Node{m, b, nd.p1, nd.depth + 1};  // ... the second half
    } else {
      st[sp++] = Node{m, b, nd.p1, nd.depth + 1};
      st[sp++] = Node{nd.p0, a, m, nd.depth + 1};
    }
  }
}

// Point of command k at ratio r: (end x, end y) and, for curves, the control point.
SWFR_HD V2 cmd_end(const CmdDev &c, double r) { return V2{lerp(c.s[0], c.e[0], r), lerp(c.s[1], c.e[1], r)}; }
SWFR_HD V2 cmd_ctrl(const CmdDev &c, double r) { return V2{lerp(c.s[2], c.e[2], r), lerp(c.s[3], c.e[3], r)}; }

// Outline of the sub-path cmds[a .. b): cmds[a] is its MoveTo (or a == 0 with an implicit start at the origin when
// the path does not begin with one); `start` = its first point.
SWFR_HD void stroke_subpath(const CmdDev *cmds, uint32_t a, uint32_t b, V2 start, bool has_move, double r, double width, Sink &sink) {
  const double w = width * 0.5;
  const uint32_t first_draw = has_move ? a + 1 : a;
  // ---- forward side ----
  Side side{true, V2{0, 0}};
  bool any = false;
  V2 t_first{0, 0}, p_first{0, 0}, t_end{0, 0}, p_end{0, 0};
  V2 cur = start;
  auto fwd = [&](const Span &s) {
    if (!any) {
      any = true;
      t_first = s.t0;
      p_first = s.p0;
    }
    offset_span(s, w, side, sink);
    t_end = s.t1;
    p_end = s.p1;
  };
  for (uint32_t k = first_draw; k < b; k++) {
    const V2 p = cmd_end(cmds[k], r);
    if (cmds[k].type == 0) {
      V2 t;
      if (unit(p.x - cur.x, p.y - cur.y, t)) fwd(Span{cur, cur, p, t, t, false});
    } else {
      quad_spans<false>(cur, cmd_ctrl(cmds[k], r), p, fwd);
    }
    cur = p;
  }
  if (!any) return;
  {  // round cap at the end
    const V2 u{t_end.y, -t_end.x};
    arc(p_end, u, t_end, w, sink);
    arc(p_end, t_end, V2{-u.x, -u.y}, w, sink);
  }
  // ---- backward side: the same spans, last to first, turned around ----
  Side back{true, V2{0, 0}};
  auto bwd = [&](const Span &s) { offset_span(s, w, back, sink); };
  for (uint32_t k = b; k-- > first_draw;) {
    const V2 p = cmd_end(cmds[k], r);
    const V2 from = k == first_draw ? start : cmd_end(cmds[k - 1], r);
    if (cmds[k].type == 0) {
      V2 t;
      if (unit(p.x - from.x, p.y - from.y, t)) bwd(Span{p, from, from, V2{-t.x, -t.y}, V2{-t.x, -t.y}, false});
    } else {
      quad_spans<true>(from, cmd_ctrl(cmds[k], r), p, bwd);
    }
  }
  {  // round cap at the start
    const V2 t_start{-t_first.x, -t_first.y};
    const V2 u{t_start.y, -t_start.x};
    arc(p_first, u, t_start, w, sink);
    arc(p_first, t_start, V2{-u.x, -u.y}, w, sink);
  }
  sink.close();
}

// Outline of one line path (all its sub-paths) at ratio r and the given width.
SWFR_HD void stroke_line(const CmdDev *cmds, uint32_t n, double r, double width, Sink &sink) {
  V2 cur{0, 0}, start{0, 0};
  uint32_t a = 0;
  int n_in_subpath = 0;
  bool has_move = false;
  for (uint32_t k = 0; k <= n; k++) {
    if (k == n || cmds[k].type == 2) {
      if (n_in_subpath > 1) stroke_subpath(cmds, a, k, start, has_move, r, width, sink);
      if (k == n) break;
      cur = cmd_end(cmds[k], r);
      start = cur;
      a = k;
      has_move = true;
      n_in_subpath = 1;
    } else {
      cur = cmd_end(cmds[k], r);
      n_in_subpath++;
    }
  }
}

}  // namespace stroke
}  // namespace swfr
