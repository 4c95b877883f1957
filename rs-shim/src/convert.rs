//! swf-tree 0.8.0 tags -> the POD descs of `include/swfr.h`.
//!
//! The C structs point at arrays (styles, colour stops, records, nested `new_styles`); `CShape` owns those arrays in
//! `Vec`s that never reallocate after `finish()`, so the pointers stay valid for as long as the `CShape` lives.  The
//! library copies what it needs during `swfr_register_*` (inputs are borrowed for the call only, like
//! `ClientAssetStore::register_shape(&mut self, tag: &DefineShape)`, rs/src/asset.rs:9-12).
//!
//! Semantic notes (each mirrors what the reference's own decoders read):
//!  * `Sfixed16P16` / `Sfixed8P8` are passed as their `epsilons` integers (ts decode-swf-shape.ts:109-125 divides by
//!    65536 / 256 itself);
//!  * `RadialGradient` is a focal gradient with focal point 0 (decode-swf-shape.ts:127-133) - the library does that,
//!    the type tag is passed through;
//!  * line caps / joins / scaling flags are not consulted by the reference renderer
//!    (canvas-renderer.ts:339-349: butt / miter defaults; :252-266: round / round for morph lines) and are dropped;
//!  * morph colour stops carry `morph_ratio`, which the reference ignores (decode-swf-morph-shape.ts:94-116 throws
//!    on gradients altogether); it is dropped here too.

use crate::b200::ffi::*;
use swf_tree::fill_styles as fs;
use swf_tree::shape_records as sr;
use swf_tree::tags::{DefineMorphShape, DefineShape};
use swf_tree::{
  ColorSpace, ColorStop, FillStyle, Gradient, GradientSpread, LineStyle, Matrix, MorphColorStop, MorphFillStyle,
  MorphGradient, MorphLineStyle, MorphShapeRecord, MorphShapeStyles, Rect, ShapeRecord, ShapeStyles, StraightSRgba8,
};

fn rgba(c: &StraightSRgba8) -> swfr_rgba8 {
  swfr_rgba8 { r: c.r, g: c.g, b: c.b, a: c.a }
}

fn matrix(m: &Matrix) -> swfr_swf_matrix {
  swfr_swf_matrix {
    scale_x: m.scale_x.epsilons,
    scale_y: m.scale_y.epsilons,
    rotate_skew0: m.rotate_skew0.epsilons,
    rotate_skew1: m.rotate_skew1.epsilons,
    translate_x: m.translate_x,
    translate_y: m.translate_y,
  }
}

fn rect(r: &Rect) -> [i32; 4] {
  [r.x_min, r.x_max, r.y_min, r.y_max]
}

fn spread(s: GradientSpread) -> u8 {
  match s {
    GradientSpread::Pad => SWFR_SPREAD_PAD,
    GradientSpread::Reflect => SWFR_SPREAD_REFLECT,
    GradientSpread::Repeat => SWFR_SPREAD_REPEAT,
  }
}

fn color_space(c: ColorSpace) -> u8 {
  match c {
    ColorSpace::SRgb => SWFR_COLOR_SRGB,
    ColorSpace::LinearRgb => SWFR_COLOR_LINEAR_RGB,
  }
}

const NO_GRADIENT: swfr_gradient = swfr_gradient { spread: 0, color_space: 0, n_colors: 0, colors: std::ptr::null() };

fn blank_fill(type_: u32) -> swfr_fill_style {
  swfr_fill_style {
    type_,
    color: swfr_rgba8::default(),
    morph_color: swfr_rgba8::default(),
    matrix: swfr_swf_matrix::default(),
    gradient: NO_GRADIENT,
    focal_point: 0,
    bitmap_id: 0,
    repeating: 0,
    smoothed: 0,
  }
}

/// Index ranges into the arenas, resolved to pointers by `finish()` (a `Vec` may move while it grows).
#[derive(Clone, Copy, Default)]
struct StylesIx {
  fill: (usize, usize),
  line: (usize, usize),
}

/// A `swfr_define_shape` together with everything it points at.
pub struct CShape {
  stops: Vec<swfr_color_stop>,
  stop_ranges: Vec<(usize, usize)>, // per fill (parallel to `fills`): its colour stops
  fills: Vec<swfr_fill_style>,
  lines: Vec<swfr_line_style>,
  line_fill_stops: Vec<(usize, usize)>, // per line: the colour stops of its fill
  styles: Vec<swfr_styles>,             // [0] = initial styles, then one per record with new_styles
  styles_ix: Vec<StylesIx>,
  records: Vec<swfr_shape_record>,
  record_styles: Vec<Option<usize>>, // per record: index into `styles`
  tag: swfr_define_shape,
}

impl CShape {
  fn new() -> Self {
    CShape {
      stops: Vec::new(),
      stop_ranges: Vec::new(),
      fills: Vec::new(),
      lines: Vec::new(),
      line_fill_stops: Vec::new(),
      styles: Vec::new(),
      styles_ix: Vec::new(),
      records: Vec::new(),
      record_styles: Vec::new(),
      tag: swfr_define_shape {
        id: 0,
        bounds: [0; 4],
        morph_bounds: [0; 4],
        initial_styles: swfr_styles { n_fill: 0, fill: std::ptr::null(), n_line: 0, line: std::ptr::null() },
        n_records: 0,
        records: std::ptr::null(),
      },
    }
  }

  pub fn as_ptr(&self) -> *const swfr_define_shape {
    &self.tag
  }

  // ---- gradients -----------------------------------------------------------------------------------------

  fn push_gradient(&mut self, g: &Gradient) -> ((usize, usize), u8, u8) {
    let first = self.stops.len();
    for ColorStop { ratio, color } in g.colors.iter() {
      self.stops.push(swfr_color_stop { ratio: *ratio, color: rgba(color), morph_color: rgba(color) });
    }
    ((first, self.stops.len()), spread(g.spread), color_space(g.color_space))
  }

  fn push_morph_gradient(&mut self, g: &MorphGradient) -> ((usize, usize), u8, u8) {
    let first = self.stops.len();
    for MorphColorStop { ratio, color, morph_color, .. } in g.colors.iter() {
      self.stops.push(swfr_color_stop { ratio: *ratio, color: rgba(color), morph_color: rgba(morph_color) });
    }
    ((first, self.stops.len()), spread(g.spread), color_space(g.color_space))
  }

  // ---- fill styles ---------------------------------------------------------------------------------------

  /// Returns the C fill and the range of its colour stops (empty for non-gradients).
  fn fill(&mut self, f: &FillStyle) -> (swfr_fill_style, (usize, usize)) {
    match f {
      FillStyle::Solid(fs::Solid { color }) => {
        let mut o = blank_fill(SWFR_FILL_SOLID);
        o.color = rgba(color);
        o.morph_color = o.color;
        (o, (0, 0))
      }
      FillStyle::Bitmap(fs::Bitmap { bitmap_id, matrix: m, repeating, smoothed }) => {
        let mut o = blank_fill(SWFR_FILL_BITMAP);
        o.bitmap_id = *bitmap_id;
        o.matrix = matrix(m);
        o.repeating = *repeating as u8;
        o.smoothed = *smoothed as u8;
        (o, (0, 0))
      }
      FillStyle::LinearGradient(fs::LinearGradient { matrix: m, gradient }) => {
        let (range, sp, cs) = self.push_gradient(gradient);
        let mut o = blank_fill(SWFR_FILL_LINEAR_GRADIENT);
        o.matrix = matrix(m);
        o.gradient.spread = sp;
        o.gradient.color_space = cs;
        (o, range)
      }
      FillStyle::RadialGradient(fs::RadialGradient { matrix: m, gradient }) => {
        let (range, sp, cs) = self.push_gradient(gradient);
        let mut o = blank_fill(SWFR_FILL_RADIAL_GRADIENT);
        o.matrix = matrix(m);
        o.gradient.spread = sp;
        o.gradient.color_space = cs;
        (o, range)
      }
      FillStyle::FocalGradient(fs::FocalGradient { matrix: m, gradient, focal_point }) => {
        let (range, sp, cs) = self.push_gradient(gradient);
        let mut o = blank_fill(SWFR_FILL_FOCAL_GRADIENT);
        o.matrix = matrix(m);
        o.gradient.spread = sp;
        o.gradient.color_space = cs;
        o.focal_point = focal_point.epsilons;
        (o, range)
      }
    }
  }

  fn morph_fill(&mut self, f: &MorphFillStyle) -> (swfr_fill_style, (usize, usize)) {
    match f {
      MorphFillStyle::Solid(fs::MorphSolid { color, morph_color }) => {
        let mut o = blank_fill(SWFR_FILL_SOLID);
        o.color = rgba(color);
        o.morph_color = rgba(morph_color);
        (o, (0, 0))
      }
      // The reference's morph decoder throws "Unknown fill type" for everything but solid fills
      // (decode-swf-morph-shape.ts:94-116); the library answers SWFR_ERR_UNSUPPORTED_STYLE for these tags, so
      // the start-state fields are enough to name the style.
      MorphFillStyle::Bitmap(fs::MorphBitmap { bitmap_id, matrix: m, repeating, smoothed, .. }) => {
        let mut o = blank_fill(SWFR_FILL_BITMAP);
        o.bitmap_id = *bitmap_id;
        o.matrix = matrix(m);
        o.repeating = *repeating as u8;
        o.smoothed = *smoothed as u8;
        (o, (0, 0))
      }
      MorphFillStyle::LinearGradient(fs::MorphLinearGradient { matrix: m, gradient, .. }) => {
        let (range, sp, cs) = self.push_morph_gradient(gradient);
        let mut o = blank_fill(SWFR_FILL_LINEAR_GRADIENT);
        o.matrix = matrix(m);
        o.gradient.spread = sp;
        o.gradient.color_space = cs;
        (o, range)
      }
      MorphFillStyle::RadialGradient(fs::MorphRadialGradient { matrix: m, gradient, .. }) => {
        let (range, sp, cs) = self.push_morph_gradient(gradient);
        let mut o = blank_fill(SWFR_FILL_RADIAL_GRADIENT);
        o.matrix = matrix(m);
        o.gradient.spread = sp;
        o.gradient.color_space = cs;
        (o, range)
      }
      MorphFillStyle::FocalGradient(fs::MorphFocalGradient { matrix: m, gradient, focal_point, .. }) => {
        let (range, sp, cs) = self.push_morph_gradient(gradient);
        let mut o = blank_fill(SWFR_FILL_FOCAL_GRADIENT);
        o.matrix = matrix(m);
        o.gradient.spread = sp;
        o.gradient.color_space = cs;
        o.focal_point = focal_point.epsilons;
        (o, range)
      }
    }
  }

  // ---- style sets ----------------------------------------------------------------------------------------

  fn push_styles(&mut self, s: &ShapeStyles) -> usize {
    let f0 = self.fills.len();
    for f in s.fill.iter() {
      let (c, range) = self.fill(f);
      self.fills.push(c);
      self.stop_ranges.push(range);
    }
    let l0 = self.lines.len();
    for LineStyle { width, fill, .. } in s.line.iter() {
      let (c, range) = self.fill(fill);
      self.lines.push(swfr_line_style { width: *width, morph_width: *width, fill: c });
      self.line_fill_stops.push(range);
    }
    self.styles_ix.push(StylesIx { fill: (f0, self.fills.len()), line: (l0, self.lines.len()) });
    self.styles_ix.len() - 1
  }

  fn push_morph_styles(&mut self, s: &MorphShapeStyles) -> usize {
    let f0 = self.fills.len();
    for f in s.fill.iter() {
      let (c, range) = self.morph_fill(f);
      self.fills.push(c);
      self.stop_ranges.push(range);
    }
    let l0 = self.lines.len();
    for MorphLineStyle { width, morph_width, fill, .. } in s.line.iter() {
      let (c, range) = self.morph_fill(fill);
      self.lines.push(swfr_line_style { width: *width, morph_width: *morph_width, fill: c });
      self.line_fill_stops.push(range);
    }
    self.styles_ix.push(StylesIx { fill: (f0, self.fills.len()), line: (l0, self.lines.len()) });
    self.styles_ix.len() - 1
  }

  // ---- records -------------------------------------------------------------------------------------------

  fn blank_record(type_: u32) -> swfr_shape_record {
    // SAFETY: swfr_shape_record is plain old data; all-zero is "no optional field present"
    let mut r: swfr_shape_record = unsafe { std::mem::zeroed() };
    r.type_ = type_;
    r
  }

  fn push_record(&mut self, r: &ShapeRecord) {
    match r {
      ShapeRecord::Edge(sr::Edge { delta, control_delta }) => {
        let mut o = Self::blank_record(SWFR_RECORD_EDGE);
        o.delta_x = delta.x;
        o.delta_y = delta.y;
        if let Some(c) = control_delta {
          o.has_control_delta = 1;
          o.control_delta_x = c.x;
          o.control_delta_y = c.y;
        }
        self.records.push(o);
        self.record_styles.push(None);
      }
      ShapeRecord::StyleChange(sr::StyleChange { move_to, left_fill, right_fill, line_style, new_styles }) => {
        let mut o = Self::blank_record(SWFR_RECORD_STYLE_CHANGE);
        if let Some(p) = move_to {
          o.has_move_to = 1;
          o.move_to_x = p.x;
          o.move_to_y = p.y;
        }
        if let Some(i) = left_fill {
          o.has_left_fill = 1;
          o.left_fill = *i as u32;
        }
        if let Some(i) = right_fill {
          o.has_right_fill = 1;
          o.right_fill = *i as u32;
        }
        if let Some(i) = line_style {
          o.has_line_style = 1;
          o.line_style = *i as u32;
        }
        let styles = new_styles.as_ref().map(|s| self.push_styles(s));
        o.has_new_styles = styles.is_some() as u8;
        self.records.push(o);
        self.record_styles.push(styles);
      }
    }
  }

  fn push_morph_record(&mut self, r: &MorphShapeRecord) {
    match r {
      MorphShapeRecord::Edge(sr::MorphEdge { delta, morph_delta, control_delta, morph_control_delta }) => {
        let mut o = Self::blank_record(SWFR_RECORD_EDGE);
        o.delta_x = delta.x;
        o.delta_y = delta.y;
        o.morph_delta_x = morph_delta.x;
        o.morph_delta_y = morph_delta.y;
        if let Some(c) = control_delta {
          o.has_control_delta = 1;
          o.control_delta_x = c.x;
          o.control_delta_y = c.y;
        }
        if let Some(c) = morph_control_delta {
          o.has_morph_control_delta = 1;
          o.morph_control_delta_x = c.x;
          o.morph_control_delta_y = c.y;
        }
        self.records.push(o);
        self.record_styles.push(None);
      }
      MorphShapeRecord::StyleChange(sr::MorphStyleChange {
        move_to,
        morph_move_to,
        left_fill,
        right_fill,
        line_style,
        new_styles,
      }) => {
        let mut o = Self::blank_record(SWFR_RECORD_STYLE_CHANGE);
        if let Some(p) = move_to {
          o.has_move_to = 1;
          o.move_to_x = p.x;
          o.move_to_y = p.y;
        }
        if let Some(p) = morph_move_to {
          o.has_morph_move_to = 1;
          o.morph_move_to_x = p.x;
          o.morph_move_to_y = p.y;
        }
        if let Some(i) = left_fill {
          o.has_left_fill = 1;
          o.left_fill = *i as u32;
        }
        if let Some(i) = right_fill {
          o.has_right_fill = 1;
          o.right_fill = *i as u32;
        }
        if let Some(i) = line_style {
          o.has_line_style = 1;
          o.line_style = *i as u32;
        }
        let styles = new_styles.as_ref().map(|s| self.push_morph_styles(s));
        o.has_new_styles = styles.is_some() as u8;
        self.records.push(o);
        self.record_styles.push(styles);
      }
    }
  }

  // ---- pointer fix-up ------------------------------------------------------------------------------------

  /// Every array has its final address now: write the pointers.
  fn finish(mut self) -> Self {
    let stops = self.stops.as_ptr();
    let at = |range: (usize, usize)| -> (*const swfr_color_stop, u16) {
      if range.1 > range.0 {
        // SAFETY: range lies inside `stops`
        (unsafe { stops.add(range.0) }, (range.1 - range.0) as u16)
      } else {
        (std::ptr::null(), 0)
      }
    };
    for (f, range) in self.fills.iter_mut().zip(self.stop_ranges.iter()) {
      let (p, n) = at(*range);
      f.gradient.colors = p;
      f.gradient.n_colors = n;
    }
    for (l, range) in self.lines.iter_mut().zip(self.line_fill_stops.iter()) {
      let (p, n) = at(*range);
      l.fill.gradient.colors = p;
      l.fill.gradient.n_colors = n;
    }
    let (fills, lines) = (self.fills.as_ptr(), self.lines.as_ptr());
    self.styles = self
      .styles_ix
      .iter()
      .map(|ix| swfr_styles {
        n_fill: (ix.fill.1 - ix.fill.0) as u32,
        // SAFETY: the ranges lie inside `fills` / `lines`
        fill: unsafe { fills.add(ix.fill.0) },
        n_line: (ix.line.1 - ix.line.0) as u32,
        line: unsafe { lines.add(ix.line.0) },
      })
      .collect();
    let styles = self.styles.as_ptr();
    for (r, s) in self.records.iter_mut().zip(self.record_styles.iter()) {
      // SAFETY: indices come from push_styles
      r.new_styles = s.map_or(std::ptr::null(), |i| unsafe { styles.add(i) });
    }
    self.tag.initial_styles = self.styles[0];
    self.tag.n_records = self.records.len() as u32;
    self.tag.records = self.records.as_ptr();
    self
  }
}

/// `&DefineShape` -> a C desc that lives as long as the returned value.
pub fn to_c_define_shape(tag: &DefineShape) -> CShape {
  let mut c = CShape::new();
  c.tag.id = tag.id;
  c.tag.bounds = rect(&tag.bounds);
  c.tag.morph_bounds = c.tag.bounds;
  c.push_styles(&tag.shape.initial_styles); // styles[0]
  for r in tag.shape.records.iter() {
    c.push_record(r);
  }
  c.finish()
}

/// `&DefineMorphShape` -> a C desc that lives as long as the returned value.
pub fn to_c_define_morph_shape(tag: &DefineMorphShape) -> CShape {
  let mut c = CShape::new();
  c.tag.id = tag.id;
  c.tag.bounds = rect(&tag.bounds);
  c.tag.morph_bounds = rect(&tag.morph_bounds);
  c.push_morph_styles(&tag.shape.initial_styles); // styles[0]
  for r in tag.shape.records.iter() {
    c.push_morph_record(r);
  }
  c.finish()
}
