// Stroke-to-fill expansion in user space (twips), run on the host when a definition is registered.
//
// Semantics: ctx.stroke() as the reference calls it
//   ts/src/lib/renderers/canvas-renderer.ts:339-349  shapes: Canvas defaults (butt caps, miter joins, limit 10)
//   ts/src/lib/renderers/canvas-renderer.ts:252-266  morph shapes: round caps and joins
// Sub-paths are never closed by the reference, so each one gets two caps.  The outline of each sub-path is one
// closed contour filled non-zero; curved pieces are offset as quadratics so the outline is resolution
// independent and is flattened on the device like any other path.  Only + - * / sqrt are used.
#include <cmath>

#include "host_types.h"

namespace swfr {
namespace {

constexpr double kMiterLimit = 10.0;
constexpr double kCosSplit = 0.9659258262890683;  // cos 15 deg
constexpr double kCosArc = 0.7071067811865476;    // cos 45 deg

struct V2 {
  double x, y;
};
struct Span {  // an offsettable piece: line (curve=false) or quadratic with small turning
  V2 p0, c, p1, t0, t1;
  bool curve;
};
struct Out {  // outline command: line to p, or quadratic (c, p)
  bool curve;
  V2 c, p;
};

bool unit(double dx, double dy, V2 &u) {
  double l = std::sqrt(dx * dx + dy * dy);
  if (l == 0.0) return false;
  u.x = dx / l;
  u.y = dy / l;
  return true;
}

void split_quad(V2 p0, V2 c, V2 p1, std::vector<Span> &out, int depth) {
  V2 t0, t1;
  bool h0 = unit(c.x - p0.x, c.y - p0.y, t0);
  bool h1 = unit(p1.x - c.x, p1.y - c.y, t1);
  if (!h0 && !h1) {
    V2 ch;
    if (unit(p1.x - p0.x, p1.y - p0.y, ch)) out.push_back(Span{p0, p0, p1, ch, ch, false});
    return;
  }
  if (!h0) t0 = t1;
  if (!h1) t1 = t0;
  double dot = t0.x * t1.x + t0.y * t1.y;
  if (dot >= kCosSplit || depth >= 8) {
    out.push_back(Span{p0, c, p1, t0, t1, true});
    return;
  }
  V2 a{(p0.x + c.x) * 0.5, (p0.y + c.y) * 0.5};
  V2 b{(c.x + p1.x) * 0.5, (c.y + p1.y) * 0.5};
  V2 m{(a.x + b.x) * 0.5, (a.y + b.y) * 0.5};
  split_quad(p0, a, m, out, depth + 1);
  split_quad(m, b, p1, out, depth + 1);
}

void arc(V2 center, V2 u, V2 v, double w, std::vector<Out> &out, int depth) {
  double dot = u.x * v.x + u.y * v.y;
  if (dot < kCosArc && depth < 6) {
    V2 mid;
    if (!unit(u.x + v.x, u.y + v.y, mid)) mid = V2{-u.y, u.x};
    arc(center, u, mid, w, out, depth + 1);
    arc(center, mid, v, w, out, depth + 1);
    return;
  }
  double k = w / (1.0 + dot);
  Out o;
  o.curve = true;
  o.c = V2{center.x + (u.x + v.x) * k, center.y + (u.y + v.y) * k};
  o.p = V2{center.x + v.x * w, center.y + v.y * w};
  out.push_back(o);
}

void line_to(std::vector<Out> &out, V2 p) { out.push_back(Out{false, p, p}); }

// Left-offset outline of consecutive spans with joins; "left" is (ty, -tx) in y-down coordinates.
void offset_side(const std::vector<Span> &spans, double w, bool round_join, std::vector<Out> &out) {
  bool first = true;
  V2 prev{0, 0};
  for (const Span &s : spans) {
    V2 n0{s.t0.y * w, -s.t0.x * w};
    V2 n1{s.t1.y * w, -s.t1.x * w};
    V2 start{s.p0.x + n0.x, s.p0.y + n0.y};
    if (first) {
      line_to(out, start);
      first = false;
    } else {
      double cross = prev.x * s.t0.y - prev.y * s.t0.x;
      double dot = prev.x * s.t0.x + prev.y * s.t0.y;
      if (dot > 0.0 && std::fabs(cross) < 1e-12) {
        line_to(out, start);
      } else if (cross > 0.0 || (cross == 0.0 && dot <= 0.0)) {  // outer side of the turn
        if (round_join) {
          arc(s.p0, V2{prev.y, -prev.x}, V2{s.t0.y, -s.t0.x}, w, out, 0);
        } else {
          if (kMiterLimit * kMiterLimit * (1.0 + dot) >= 2.0) {
            double k = w / (1.0 + dot);
            line_to(out, V2{s.p0.x + (prev.y + s.t0.y) * k, s.p0.y + (-prev.x - s.t0.x) * k});
          }
          line_to(out, start);
        }
      } else {  // inner side: through the vertex, keeps the contour's winding consistent
        line_to(out, s.p0);
        line_to(out, start);
      }
    }
    V2 end{s.p1.x + n1.x, s.p1.y + n1.y};
    if (!s.curve) {
      line_to(out, end);
    } else {
      double dotn = s.t0.x * s.t1.x + s.t0.y * s.t1.y;
      double k = w / (1.0 + dotn);
      Out o;
      o.curve = true;
      o.c = V2{s.c.x + (s.t0.y + s.t1.y) * k, s.c.y + (-s.t0.x - s.t1.x) * k};
      o.p = end;
      out.push_back(o);
    }
    prev = s.t1;
  }
}

void stroke_subpath(const std::vector<Span> &spans, double width, bool round_style, std::vector<StrokeSeg> &segs) {
  if (spans.empty()) return;
  double w = width * 0.5;
  std::vector<Out> out;
  offset_side(spans, w, round_style, out);
  V2 t_end = spans.back().t1, p_end = spans.back().p1;
  if (round_style) {
    V2 u{t_end.y, -t_end.x};
    arc(p_end, u, t_end, w, out, 0);
    arc(p_end, t_end, V2{-u.x, -u.y}, w, out, 0);
  }
  std::vector<Span> back;
  for (size_t i = spans.size(); i-- > 0;) {
    const Span &s = spans[i];
    back.push_back(Span{s.p1, s.c, s.p0, V2{-s.t1.x, -s.t1.y}, V2{-s.t0.x, -s.t0.y}, s.curve});
  }
  offset_side(back, w, round_style, out);
  V2 t_start = back.back().t1, p_start = back.back().p1;
  if (round_style) {
    V2 u{t_start.y, -t_start.x};
    arc(p_start, u, t_start, w, out, 0);
    arc(p_start, t_start, V2{-u.x, -u.y}, w, out, 0);
  }
  // out[0] is the contour's first point; close back to it.  Coordinates are rounded to float32 here.
  float fx = (float)out[0].p.x, fy = (float)out[0].p.y;
  float cx = fx, cy = fy;
  auto emit_line = [&](float x, float y) {
    if (x != cx || y != cy) {
      StrokeSeg s;
      s.curve = false;
      s.p[0] = cx, s.p[1] = cy, s.p[2] = cx, s.p[3] = cy, s.p[4] = x, s.p[5] = y;
      segs.push_back(s);
    }
    cx = x, cy = y;
  };
  for (size_t i = 1; i < out.size(); i++) {
    const Out &o = out[i];
    if (!o.curve) {
      emit_line((float)o.p.x, (float)o.p.y);
    } else {
      StrokeSeg s;
      s.curve = true;
      s.p[0] = cx, s.p[1] = cy, s.p[2] = (float)o.c.x, s.p[3] = (float)o.c.y, s.p[4] = (float)o.p.x, s.p[5] = (float)o.p.y;
      segs.push_back(s);
      cx = s.p[4], cy = s.p[5];
    }
  }
  emit_line(fx, fy);
}

}  // namespace

void stroke_commands(const std::vector<Command> &cmds, double width, bool round_style, std::vector<StrokeSeg> &out) {
  std::vector<Span> spans;
  V2 cur{0, 0};
  int n_in_subpath = 0;
  auto flush = [&]() {
    if (n_in_subpath > 1) stroke_subpath(spans, width, round_style, out);
    spans.clear();
  };
  for (const Command &c : cmds) {
    if (c.type == 2) {
      flush();
      cur = V2{c.s[0], c.s[1]};
      n_in_subpath = 1;
    } else if (c.type == 0) {
      V2 p{c.s[0], c.s[1]}, t;
      if (unit(p.x - cur.x, p.y - cur.y, t)) spans.push_back(Span{cur, cur, p, t, t, false});
      cur = p;
      n_in_subpath++;
    } else {
      V2 cp{c.s[2], c.s[3]}, p{c.s[0], c.s[1]};
      split_quad(cur, cp, p, spans, 0);
      cur = p;
      n_in_subpath++;
    }
  }
  flush();
}

}  // namespace swfr
