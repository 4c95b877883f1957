// Host pre-pass run once per definition (where the reference caches it, canvas-renderer.ts:96-112):
// shape records -> ordered style paths -> device segments.
//
// Behaviour follows the reference compilers
//   ts/src/lib/shape/decode-swf-shape.ts:298-448        (style layers, left fill forward / right fill reversed)
//   ts/src/lib/shape/decode-swf-shape.ts:203-234        (single-pass chain extraction)
//   ts/src/lib/shape/decode-swf-morph-shape.ts:304-364  (pairs of coordinates, control point = delta / 2 when absent)
// and must reproduce tests/*/shape.ts.json exactly (checked through swfr_debug_compiled).
#include <cmath>
#include <cstring>
#include <deque>

#include "host_types.h"

namespace swfr {
namespace {

struct Pt2 {  // a coordinate in the start and the end state
  double v[2];
};

struct Piece {  // one edge record as seen by one style
  Pt2 sx, sy, cx, cy, ex, ey;
  bool curved;
};

struct StyleSet {
  int style_index;  // index into the layer's fill or line styles
  std::vector<Piece> pieces;
};

struct Layer {
  std::vector<swfr_fill_style> fill_styles;
  std::vector<swfr_line_style> line_styles;
  std::vector<StyleSet> fills, lines;
};

Layer make_layer(const swfr_styles &st) {
  Layer l;
  for (uint32_t i = 0; i < st.n_fill; i++) {
    l.fill_styles.push_back(st.fill[i]);
    l.fills.push_back(StyleSet{(int)i, {}});
  }
  for (uint32_t i = 0; i < st.n_line; i++) {
    l.line_styles.push_back(st.line[i]);
    l.lines.push_back(StyleSet{(int)i, {}});
  }
  return l;
}

// Pulls one chain out of `open`: the first piece, then ONE forward sweep over the rest, appending a piece
// whose start meets the chain's tail or prepending one whose end meets the chain's head.  Only the start
// state is compared (decode-swf-morph-shape.ts:176-196).
std::deque<Piece> take_chain(std::vector<Piece> &open) {
  std::deque<Piece> chain;
  chain.push_back(open.front());
  open.erase(open.begin());
  double hx = chain.front().sx.v[0], hy = chain.front().sy.v[0];
  double tx = chain.front().ex.v[0], ty = chain.front().ey.v[0];
  for (size_t i = 0; i < open.size();) {
    const Piece &c = open[i];
    if (c.sx.v[0] == tx && c.sy.v[0] == ty) {
      tx = c.ex.v[0];
      ty = c.ey.v[0];
      chain.push_back(c);
      open.erase(open.begin() + i);
    } else if (c.ex.v[0] == hx && c.ey.v[0] == hy) {
      hx = c.sx.v[0];
      hy = c.sy.v[0];
      chain.push_front(c);
      open.erase(open.begin() + i);
    } else {
      i++;
    }
  }
  return chain;
}

void pieces_to_commands(const std::vector<Piece> &pieces, std::vector<Command> &out) {
  std::vector<Piece> open(pieces);
  while (!open.empty()) {
    std::deque<Piece> chain = take_chain(open);
    Command mv{};
    mv.type = 2;
    mv.s[0] = chain.front().sx.v[0], mv.s[1] = chain.front().sy.v[0];
    mv.e[0] = chain.front().sx.v[1], mv.e[1] = chain.front().sy.v[1];
    out.push_back(mv);
    for (const Piece &p : chain) {
      Command c{};
      c.type = p.curved ? 1 : 0;
      c.s[0] = p.ex.v[0], c.s[1] = p.ey.v[0];
      c.e[0] = p.ex.v[1], c.e[1] = p.ey.v[1];
      if (p.curved) {
        c.s[2] = p.cx.v[0], c.s[3] = p.cy.v[0];
        c.e[2] = p.cx.v[1], c.e[3] = p.cy.v[1];
      }
      out.push_back(c);
    }
  }
}

bool validate_fill(const swfr_fill_style &f, bool morph, std::string &err) {
  if (f.type > SWFR_FILL_BITMAP) {
    err = "UnknownFillStyle";
    return false;
  }
  if (morph && f.type != SWFR_FILL_SOLID) {
    err = "Unknown fill type (morph shapes support solid fills only, decode-swf-morph-shape.ts:94-106)";
    return false;
  }
  return true;
}

DefPaint paint_from_fill(const swfr_fill_style &f, bool morph, CompiledDef &def) {
  DefPaint p{};
  p.lut = -1;
  switch (f.type) {
    case SWFR_FILL_SOLID:
      p.type = PAINT_SOLID;
      memcpy(p.color0, &f.color, 4);
      memcpy(p.color1, morph ? &f.morph_color : &f.color, 4);
      if (morph) p.flags |= PF_COLOR_MORPH;
      break;
    case SWFR_FILL_BITMAP:
      p.type = PAINT_BITMAP;
      p.bitmap_id = f.bitmap_id;
      p.repeating = f.repeating ? 1 : 0;
      break;
    default: {  // gradients; radial == focal with focal point 0 (decode-swf-shape.ts:127-133)
      p.type = f.type == SWFR_FILL_LINEAR_GRADIENT ? PAINT_LINEAR : PAINT_FOCAL;
      p.spread = f.gradient.spread;
      p.focal = f.type == SWFR_FILL_FOCAL_GRADIENT ? (double)f.focal_point / 256.0 : 0.0;
      std::vector<uint32_t> ramp;
      bool opaque = false;
      build_ramp(f.gradient.colors, f.gradient.n_colors, f.gradient.color_space == SWFR_COLOR_LINEAR_RGB, false, ramp,
                 &opaque);
      if (opaque) p.flags |= PF_OPAQUE_RAMP;
      p.lut = (int32_t)def.luts.size();
      def.luts.push_back(std::move(ramp));
    }
  }
  if (f.type != SWFR_FILL_SOLID) {
    p.matrix[0] = (double)f.matrix.scale_x / 65536.0;
    p.matrix[1] = (double)f.matrix.rotate_skew0 / 65536.0;
    p.matrix[2] = (double)f.matrix.rotate_skew1 / 65536.0;
    p.matrix[3] = (double)f.matrix.scale_y / 65536.0;
    p.matrix[4] = (double)f.matrix.translate_x;
    p.matrix[5] = (double)f.matrix.translate_y;
  }
  return p;
}

void push_seg(CompiledDef &def, uint32_t path, bool curve, const double s[6], const double e[6]) {
  SegMorph g{};
  for (int i = 0; i < 6; i++) {
    g.s[i] = (float)s[i];
    g.e[i] = (float)e[i];
  }
  g.path_flags = path | (curve ? 0x80000000u : 0u);
  def.segs.push_back(g);
}

// ctx.fill() closes every sub-path implicitly (Canvas); ctx.stroke() does not.
void fill_path_to_segments(const std::vector<Command> &cmds, uint32_t path, CompiledDef &def) {
  bool have = false;
  double start_s[2] = {0, 0}, start_e[2] = {0, 0}, cur_s[2] = {0, 0}, cur_e[2] = {0, 0};
  auto close = [&]() {
    if (!have) return;
    if (cur_s[0] != start_s[0] || cur_s[1] != start_s[1] || cur_e[0] != start_e[0] || cur_e[1] != start_e[1]) {
      double s[6] = {cur_s[0], cur_s[1], cur_s[0], cur_s[1], start_s[0], start_s[1]};
      double e[6] = {cur_e[0], cur_e[1], cur_e[0], cur_e[1], start_e[0], start_e[1]};
      push_seg(def, path, false, s, e);
    }
  };
  for (const Command &c : cmds) {
    if (c.type == 2) {
      close();
      have = true;
      start_s[0] = cur_s[0] = c.s[0], start_s[1] = cur_s[1] = c.s[1];
      start_e[0] = cur_e[0] = c.e[0], start_e[1] = cur_e[1] = c.e[1];
    } else if (c.type == 0) {
      double s[6] = {cur_s[0], cur_s[1], cur_s[0], cur_s[1], c.s[0], c.s[1]};
      double e[6] = {cur_e[0], cur_e[1], cur_e[0], cur_e[1], c.e[0], c.e[1]};
      push_seg(def, path, false, s, e);
      cur_s[0] = c.s[0], cur_s[1] = c.s[1], cur_e[0] = c.e[0], cur_e[1] = c.e[1];
    } else {
      double s[6] = {cur_s[0], cur_s[1], c.s[2], c.s[3], c.s[0], c.s[1]};
      double e[6] = {cur_e[0], cur_e[1], c.e[2], c.e[3], c.e[0], c.e[1]};
      push_seg(def, path, true, s, e);
      cur_s[0] = c.s[0], cur_s[1] = c.s[1], cur_e[0] = c.e[0], cur_e[1] = c.e[1];
    }
  }
  close();
}

}  // namespace

int compile_definition(const swfr_define_shape *tag, bool morph, CompiledDef &out, std::string &err) {
  out = CompiledDef{};
  out.is_morph = morph;
  std::vector<Layer> layers;
  StyleSet *left = nullptr, *right = nullptr, *line = nullptr;
  int left_i = -1, right_i = -1, line_i = -1;  // indices survive vector growth, pointers are refreshed below
  auto refresh = [&]() {
    Layer &l = layers.back();
    left = left_i >= 0 ? &l.fills[left_i] : nullptr;
    right = right_i >= 0 ? &l.fills[right_i] : nullptr;
    line = line_i >= 0 ? &l.lines[line_i] : nullptr;
  };
  auto new_layer = [&](const swfr_styles &st) {
    layers.push_back(make_layer(st));
    left_i = right_i = line_i = -1;  // a new style list resets the three selections (decode-swf-shape.ts:402-408)
    refresh();
  };
  auto select = [&](uint32_t id, size_t n, int &idx) -> bool {
    if (id == 0) {
      idx = -1;
      return true;
    }
    if (id - 1 >= n) return false;
    idx = (int)(id - 1);
    return true;
  };
  for (uint32_t i = 0; i < tag->initial_styles.n_fill; i++)
    if (!validate_fill(tag->initial_styles.fill[i], morph, err)) return SWFR_ERR_UNSUPPORTED_STYLE;
  for (uint32_t i = 0; i < tag->initial_styles.n_line; i++)
    if (!validate_fill(tag->initial_styles.line[i].fill, morph, err)) return SWFR_ERR_UNSUPPORTED_STYLE;
  new_layer(tag->initial_styles);

  Pt2 x{{0, 0}}, y{{0, 0}};
  for (uint32_t ri = 0; ri < tag->n_records; ri++) {
    const swfr_shape_record &rec = tag->records[ri];
    if (rec.type == SWFR_RECORD_STYLE_CHANGE) {
      // order matters: new styles, left, right, line, move (decode-swf-shape.ts:337-356); the morph decoder has
      // no new-styles handling (decode-swf-morph-shape.ts:304-322)
      if (!morph && rec.has_new_styles && rec.new_styles) {
        for (uint32_t i = 0; i < rec.new_styles->n_fill; i++)
          if (!validate_fill(rec.new_styles->fill[i], morph, err)) return SWFR_ERR_UNSUPPORTED_STYLE;
        for (uint32_t i = 0; i < rec.new_styles->n_line; i++)
          if (!validate_fill(rec.new_styles->line[i].fill, morph, err)) return SWFR_ERR_UNSUPPORTED_STYLE;
        new_layer(*rec.new_styles);
      }
      Layer &l = layers.back();
      if (rec.has_left_fill && !select(rec.left_fill, l.fills.size(), left_i)) {
        err = "Invalid fill ID";
        return SWFR_ERR_INVALID_FILL_ID;
      }
      if (rec.has_right_fill && !select(rec.right_fill, l.fills.size(), right_i)) {
        err = "Invalid fill ID";
        return SWFR_ERR_INVALID_FILL_ID;
      }
      if (rec.has_line_style && !select(rec.line_style, l.lines.size(), line_i)) {
        err = "Invalid fill ID";
        return SWFR_ERR_INVALID_FILL_ID;
      }
      refresh();
      if (rec.has_move_to) {
        if (morph && !rec.has_morph_move_to) {
          err = "Expected morphMoveTo to be defined";
          return SWFR_ERR_MALFORMED;
        }
        x.v[0] = rec.move_to_x, y.v[0] = rec.move_to_y;
        x.v[1] = morph ? rec.morph_move_to_x : rec.move_to_x;
        y.v[1] = morph ? rec.morph_move_to_y : rec.move_to_y;
      }
    } else if (rec.type == SWFR_RECORD_EDGE) {
      Piece p{};
      double mdx = morph ? rec.morph_delta_x : rec.delta_x, mdy = morph ? rec.morph_delta_y : rec.delta_y;
      p.sx = x, p.sy = y;
      p.ex.v[0] = x.v[0] + rec.delta_x, p.ex.v[1] = x.v[1] + mdx;
      p.ey.v[0] = y.v[0] + rec.delta_y, p.ey.v[1] = y.v[1] + mdy;
      bool has_c = rec.has_control_delta, has_mc = morph && rec.has_morph_control_delta;
      p.curved = has_c || has_mc;
      if (p.curved) {
        // a missing control point is synthesised at half the delta (decode-swf-morph-shape.ts:341-346)
        double cdx = has_c ? (double)rec.control_delta_x : rec.delta_x / 2.0;
        double cdy = has_c ? (double)rec.control_delta_y : rec.delta_y / 2.0;
        double mcdx = morph ? (has_mc ? (double)rec.morph_control_delta_x : mdx / 2.0) : cdx;
        double mcdy = morph ? (has_mc ? (double)rec.morph_control_delta_y : mdy / 2.0) : cdy;
        p.cx.v[0] = x.v[0] + cdx, p.cx.v[1] = x.v[1] + mcdx;
        p.cy.v[0] = y.v[0] + cdy, p.cy.v[1] = y.v[1] + mcdy;
      }
      if (left) left->pieces.push_back(p);
      if (right) {
        Piece q = p;  // the right fill sees the edge reversed, same control point
        q.sx = p.ex, q.sy = p.ey, q.ex = p.sx, q.ey = p.sy;
        right->pieces.push_back(q);
      }
      if (line) line->pieces.push_back(p);
      x = p.ex, y = p.ey;
    } else {
      err = "UnreachableCode: unknown record type";
      return SWFR_ERR_MALFORMED;
    }
  }

  // paint order: layer by layer, fills by style index then lines by style index; unused styles give no path
  for (Layer &l : layers) {
    for (StyleSet &fs : l.fills) {
      CompiledPath cp;
      pieces_to_commands(fs.pieces, cp.commands);
      if (cp.commands.empty()) continue;
      cp.has_fill = true;
      cp.fill = l.fill_styles[fs.style_index];
      if (cp.fill.gradient.n_colors) {
        cp.stops.assign(cp.fill.gradient.colors, cp.fill.gradient.colors + cp.fill.gradient.n_colors);
      }
      out.paths.push_back(std::move(cp));
    }
    for (StyleSet &ls : l.lines) {
      CompiledPath cp;
      pieces_to_commands(ls.pieces, cp.commands);
      if (cp.commands.empty()) continue;
      cp.has_line = true;
      cp.line = l.line_styles[ls.style_index];
      out.paths.push_back(std::move(cp));
    }
  }
  for (CompiledPath &cp : out.paths) {  // re-point gradient stops at the owned copies
    if (cp.has_fill && !cp.stops.empty()) cp.fill.gradient.colors = cp.stops.data();
  }

  // device form
  double width_state = 1.0;  // Canvas default lineWidth; zero widths are ignored and the previous one stays
  for (CompiledPath &cp : out.paths) {
    if (cp.has_fill) {
      uint32_t path = (uint32_t)out.paints.size();
      out.paints.push_back(paint_from_fill(cp.fill, morph, out));
      fill_path_to_segments(cp.commands, path, out);
    }
    if (cp.has_line) {
      if (cp.line.fill.type != SWFR_FILL_SOLID) {
        err = "NotImplementedLineStyle";
        return SWFR_ERR_UNSUPPORTED_STYLE;
      }
      if (morph) {
        // stroke geometry depends on the ratio; it is expanded per draw (see renderer).  A stroke whose
        // colour is fully transparent in both states composites nothing.
        if (cp.line.fill.color.a != 0 || cp.line.fill.morph_color.a != 0) out.has_visible_morph_stroke = true;
        MorphLine ml;
        ml.commands = cp.commands;
        ml.w0 = (double)cp.line.width;
        ml.w1 = (double)cp.line.morph_width;
        memcpy(ml.color0, &cp.line.fill.color, 4);
        memcpy(ml.color1, &cp.line.fill.morph_color, 4);
        out.morph_lines.push_back(std::move(ml));
        continue;
      }
      if (cp.line.width > 0) width_state = (double)cp.line.width;
      std::vector<StrokeSeg> ss;
      stroke_commands(cp.commands, width_state, false, ss);
      uint32_t path = (uint32_t)out.paints.size();
      out.paints.push_back(paint_from_fill(cp.line.fill, false, out));
      out.paints.back().flags |= PF_SAMPLED;
      for (const StrokeSeg &s : ss) {
        SegMorph g{};
        for (int i = 0; i < 6; i++) g.s[i] = g.e[i] = s.p[i];
        g.path_flags = path | (s.curve ? 0x80000000u : 0u);
        out.segs.push_back(g);
      }
    }
  }
  // bounds of every device path (start and end state: SegMorph holds s[6] then e[6] = 6 points)
  if (!out.segs.empty())
    set_paint_bounds(out.paints.data(), out.paints.size(), out.segs[0].s, out.segs.size(), sizeof(SegMorph) / sizeof(float), 6,
                     &out.segs[0].path_flags, sizeof(SegMorph));
  else
    set_paint_bounds(out.paints.data(), out.paints.size(), nullptr, 0, 0, 0, nullptr, 0);
  return SWFR_OK;
}

}  // namespace swfr
