// node-addon-api binding of libswfr_b200.so for the reference's TypeScript package (not compiled in this repository:
// no Node toolchain in the image).  One class, `Native`, used by b200-renderer.ts:
//
//   new Native(width, height, device)                 swfr_create
//   registerShape(tagJson, morph) -> id               swfr_register_shape / swfr_register_morph_shape
//   addBitmap(id, bytes)                              swfr_register_bitmap_xswfbmp        (node-canvas-bitmap-service.ts:14-37)
//   render(width, height, background?, nodes, nRoot)  swfr_render_display_stage           (canvas-renderer.ts:61-145)
//   readImage(premultiplied) -> Uint8ClampedArray     swfr_read_image                     (getImageData)
//   toPng() / toPam() -> Uint8Array                   swfr_write_png / swfr_write_pam     (spec.ts:134-147, image-data-to-pam.ts)
//   close()                                           swfr_destroy
//
// `tagJson` is swf-tree's JSON form of DefineShape / DefineMorphShape (snake_case, the corpus' ast.json); the conversion
// below is swf_renderer_b200/swf_tree.py field by field.  `nodes` is the display tree packed by B200Renderer.packStage:
// twelve int32 per node, children contiguous.
#include <napi.h>

#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "swfr.h"

namespace {

// Everything a converted tag points at lives here until the registration call has returned.
struct TagArena {
  std::deque<std::vector<swfr_color_stop>> stops;
  std::deque<std::vector<swfr_fill_style>> fills;
  std::deque<std::vector<swfr_line_style>> lines;
  std::deque<swfr_styles> styles;
  std::vector<swfr_shape_record> records;
};

bool present(const Napi::Object &o, const char *key) {
  if (!o.Has(key)) return false;
  Napi::Value v = o.Get(key);
  return !v.IsUndefined() && !v.IsNull();
}
int32_t i32(const Napi::Object &o, const char *key) { return o.Get(key).As<Napi::Number>().Int32Value(); }
uint32_t u32(const Napi::Object &o, const char *key) { return o.Get(key).As<Napi::Number>().Uint32Value(); }
std::string str(const Napi::Object &o, const char *key) { return o.Get(key).As<Napi::String>().Utf8Value(); }
Napi::Object obj(const Napi::Object &o, const char *key) { return o.Get(key).As<Napi::Object>(); }
Napi::Array arr(const Napi::Object &o, const char *key) { return o.Get(key).As<Napi::Array>(); }

swfr_rgba8 rgba(const Napi::Object &c) {
  return swfr_rgba8{(uint8_t)u32(c, "r"), (uint8_t)u32(c, "g"), (uint8_t)u32(c, "b"), (uint8_t)u32(c, "a")};
}

swfr_swf_matrix matrix(const Napi::Object &m) {
  return swfr_swf_matrix{i32(m, "scale_x"),      i32(m, "scale_y"),     i32(m, "rotate_skew0"),
                         i32(m, "rotate_skew1"), i32(m, "translate_x"), i32(m, "translate_y")};
}

swfr_fill_style fill_style(const Napi::Object &s, TagArena &arena) {
  swfr_fill_style f;
  memset(&f, 0, sizeof f);
  const std::string t = str(s, "type");
  if (t == "solid") {
    f.type = SWFR_FILL_SOLID;
    f.color = rgba(obj(s, "color"));
    f.morph_color = present(s, "morph_color") ? rgba(obj(s, "morph_color")) : f.color;
  } else if (t == "bitmap") {
    f.type = SWFR_FILL_BITMAP;
    f.bitmap_id = (uint16_t)u32(s, "bitmap_id");
    f.matrix = matrix(obj(s, "matrix"));
    f.repeating = s.Get("repeating").ToBoolean().Value() ? 1 : 0;
    f.smoothed = s.Get("smoothed").ToBoolean().Value() ? 1 : 0;
  } else if (t == "linear-gradient" || t == "radial-gradient" || t == "focal-gradient") {
    f.type = t == "linear-gradient" ? SWFR_FILL_LINEAR_GRADIENT : (t == "radial-gradient" ? SWFR_FILL_RADIAL_GRADIENT : SWFR_FILL_FOCAL_GRADIENT);
    f.matrix = matrix(obj(s, "matrix"));
    const Napi::Object g = obj(s, "gradient");
    const Napi::Array colors = arr(g, "colors");
    arena.stops.emplace_back(colors.Length());
    std::vector<swfr_color_stop> &stops = arena.stops.back();
    for (uint32_t i = 0; i < colors.Length(); i++) {
      const Napi::Object st = colors.Get(i).As<Napi::Object>();
      stops[i].ratio = (uint8_t)u32(st, "ratio");
      stops[i].color = rgba(obj(st, "color"));
      stops[i].morph_color = present(st, "morph_color") ? rgba(obj(st, "morph_color")) : stops[i].color;
    }
    const std::string spread = str(g, "spread"), space = str(g, "color_space");
    f.gradient.spread = spread == "reflect" ? SWFR_SPREAD_REFLECT : (spread == "repeat" ? SWFR_SPREAD_REPEAT : SWFR_SPREAD_PAD);
    f.gradient.color_space = space == "linear-rgb" ? SWFR_COLOR_LINEAR_RGB : SWFR_COLOR_SRGB;
    f.gradient.n_colors = (uint16_t)stops.size();
    f.gradient.colors = stops.data();
    if (t == "focal-gradient") f.focal_point = (int16_t)i32(s, "focal_point");
  } else {
    f.type = 255;  // the library reports UnknownFillStyle
  }
  return f;
}

swfr_styles styles(const Napi::Object &st, TagArena &arena) {
  const Napi::Array fill = arr(st, "fill"), line = arr(st, "line");
  arena.fills.emplace_back(fill.Length());
  arena.lines.emplace_back(line.Length());
  std::vector<swfr_fill_style> &fills = arena.fills.back();
  std::vector<swfr_line_style> &lines = arena.lines.back();
  for (uint32_t i = 0; i < fill.Length(); i++) fills[i] = fill_style(fill.Get(i).As<Napi::Object>(), arena);
  for (uint32_t i = 0; i < line.Length(); i++) {
    const Napi::Object s = line.Get(i).As<Napi::Object>();
    memset(&lines[i], 0, sizeof lines[i]);
    lines[i].width = (uint16_t)u32(s, "width");
    lines[i].morph_width = present(s, "morph_width") ? (uint16_t)u32(s, "morph_width") : lines[i].width;
    lines[i].fill = fill_style(obj(s, "fill"), arena);
  }
  swfr_styles out;
  out.n_fill = (uint32_t)fills.size();
  out.fill = fills.data();
  out.n_line = (uint32_t)lines.size();
  out.line = lines.data();
  return out;
}

void bounds(const Napi::Object &b, int32_t out[4]) {
  out[0] = i32(b, "x_min");
  out[1] = i32(b, "x_max");
  out[2] = i32(b, "y_min");
  out[3] = i32(b, "y_max");
}

// define-shape / define-morph-shape JSON -> swfr_define_shape (pointers into `arena`)
swfr_define_shape define_shape(const Napi::Object &tag, TagArena &arena) {
  swfr_define_shape t;
  memset(&t, 0, sizeof t);
  t.id = present(tag, "id") ? (uint16_t)u32(tag, "id") : 0;
  bounds(obj(tag, "bounds"), t.bounds);
  bounds(present(tag, "morph_bounds") ? obj(tag, "morph_bounds") : obj(tag, "bounds"), t.morph_bounds);
  const Napi::Object shape = obj(tag, "shape");
  t.initial_styles = styles(obj(shape, "initial_styles"), arena);
  const Napi::Array recs = arr(shape, "records");
  arena.records.resize(recs.Length());
  for (uint32_t i = 0; i < recs.Length(); i++) {
    const Napi::Object r = recs.Get(i).As<Napi::Object>();
    swfr_shape_record &o = arena.records[i];
    memset(&o, 0, sizeof o);
    const std::string type = str(r, "type");
    if (type == "edge") {
      o.type = SWFR_RECORD_EDGE;
      const Napi::Object d = obj(r, "delta");
      o.delta_x = i32(d, "x");
      o.delta_y = i32(d, "y");
      const Napi::Object md = present(r, "morph_delta") ? obj(r, "morph_delta") : d;
      o.morph_delta_x = i32(md, "x");
      o.morph_delta_y = i32(md, "y");
      if (present(r, "control_delta")) {
        const Napi::Object c = obj(r, "control_delta");
        o.has_control_delta = 1;
        o.control_delta_x = i32(c, "x");
        o.control_delta_y = i32(c, "y");
      }
      if (present(r, "morph_control_delta")) {
        const Napi::Object c = obj(r, "morph_control_delta");
        o.has_morph_control_delta = 1;
        o.morph_control_delta_x = i32(c, "x");
        o.morph_control_delta_y = i32(c, "y");
      }
    } else if (type == "style-change") {
      o.type = SWFR_RECORD_STYLE_CHANGE;
      if (present(r, "move_to")) {
        const Napi::Object p = obj(r, "move_to");
        o.has_move_to = 1;
        o.move_to_x = i32(p, "x");
        o.move_to_y = i32(p, "y");
      }
      if (present(r, "morph_move_to")) {
        const Napi::Object p = obj(r, "morph_move_to");
        o.has_morph_move_to = 1;
        o.morph_move_to_x = i32(p, "x");
        o.morph_move_to_y = i32(p, "y");
      }
      if (present(r, "left_fill")) o.has_left_fill = 1, o.left_fill = u32(r, "left_fill");
      if (present(r, "right_fill")) o.has_right_fill = 1, o.right_fill = u32(r, "right_fill");
      if (present(r, "line_style")) o.has_line_style = 1, o.line_style = u32(r, "line_style");
      if (present(r, "new_styles")) {
        arena.styles.push_back(styles(obj(r, "new_styles"), arena));
        o.has_new_styles = 1;
        o.new_styles = &arena.styles.back();  // std::deque: stable addresses
      }
    } else {
      o.type = 255;  // the library reports the unknown record type
    }
  }
  t.n_records = (uint32_t)arena.records.size();
  t.records = arena.records.data();
  return t;
}

constexpr size_t kNodeInts = 12;  // b200-renderer.ts: NODE_INTS

class Native : public Napi::ObjectWrap<Native> {
 public:
  static Napi::Object Init(Napi::Env env, Napi::Object exports) {
    Napi::Function ctor = DefineClass(env, "Native",
                                      {InstanceMethod("registerShape", &Native::RegisterShape), InstanceMethod("addBitmap", &Native::AddBitmap),
                                       InstanceMethod("render", &Native::Render), InstanceMethod("readImage", &Native::ReadImage),
                                       InstanceMethod("toPng", &Native::ToPng), InstanceMethod("toPam", &Native::ToPam),
                                       InstanceMethod("close", &Native::Close)});
    exports.Set("Native", ctor);
    return exports;
  }

  explicit Native(const Napi::CallbackInfo &info) : Napi::ObjectWrap<Native>(info) {
    width_ = info[0].As<Napi::Number>().Uint32Value();
    height_ = info[1].As<Napi::Number>().Uint32Value();
    const int device = info.Length() > 2 ? info[2].As<Napi::Number>().Int32Value() : 0;
    // no CUDA device -> SWFR_ERR_CUDA: there is no CPU fallback behind this binding
    const int rc = swfr_create(device, width_, height_, &r_);
    if (rc != SWFR_OK) throw Napi::Error::New(info.Env(), std::string("swfr_create: ") + swfr_status_string(rc));
  }
  ~Native() override {
    if (r_) swfr_destroy(r_);
  }

 private:
  void check(Napi::Env env, int rc, const char *what) {
    if (rc == SWFR_OK) return;
    const char *detail = r_ ? swfr_last_error(r_) : "";
    // the reference's error names travel in the detail text (UnknownFillStyle, BitmapNotFound, ...: INTEGRATION.md section 4)
    throw Napi::Error::New(env, std::string(what) + ": " + swfr_status_string(rc) + (detail && *detail ? std::string(": ") + detail : ""));
  }

  Napi::Value RegisterShape(const Napi::CallbackInfo &info) {
    TagArena arena;
    const swfr_define_shape tag = define_shape(info[0].As<Napi::Object>(), arena);
    const bool morph = info[1].ToBoolean().Value();
    uint32_t id = 0;
    check(info.Env(), morph ? swfr_register_morph_shape(r_, &tag, &id) : swfr_register_shape(r_, &tag, &id), "registerShape");
    return Napi::Number::New(info.Env(), id);
  }

  Napi::Value AddBitmap(const Napi::CallbackInfo &info) {
    const uint32_t id = info[0].As<Napi::Number>().Uint32Value();
    const Napi::Uint8Array data = info[1].As<Napi::Uint8Array>();
    check(info.Env(), swfr_register_bitmap_xswfbmp(r_, (uint16_t)id, data.Data(), data.ByteLength()), "addBitmap");
    return info.Env().Undefined();
  }

  // nodes: [type, id, hasMatrix, scaleX, scaleY, rotateSkew0, rotateSkew1, translateX, translateY, ratio bits, childCount,
  // firstChild] per node; nodes 0 .. nRoot - 1 are the stage's children
  Napi::Value Render(const Napi::CallbackInfo &info) {
    swfr_display_stage stage;
    memset(&stage, 0, sizeof stage);
    stage.width = info[0].As<Napi::Number>().Uint32Value();
    stage.height = info[1].As<Napi::Number>().Uint32Value();
    if (!info[2].IsUndefined() && !info[2].IsNull()) {
      const Napi::Uint8Array bg = info[2].As<Napi::Uint8Array>();
      stage.has_background_color = 1;
      stage.background_color = swfr_rgba8{bg[0], bg[1], bg[2], bg[3]};
    }
    const Napi::Int32Array nodes = info[3].As<Napi::Int32Array>();
    const size_t n = nodes.ElementLength() / kNodeInts;
    const uint32_t n_root = info[4].As<Napi::Number>().Uint32Value();
    if (n_root > n) throw Napi::Error::New(info.Env(), "render: packed display tree out of range");
    std::vector<swfr_display_object> objs(n);
    for (size_t i = 0; i < n; i++) {
      const int32_t *v = nodes.Data() + i * kNodeInts;
      swfr_display_object &o = objs[i];
      memset(&o, 0, sizeof o);
      o.type = (uint32_t)v[0];
      o.id = (uint32_t)v[1];
      o.has_matrix = v[2] ? 1 : 0;
      o.matrix = swfr_swf_matrix{v[3], v[4], v[5], v[6], v[7], v[8]};
      memcpy(&o.ratio, &v[9], sizeof(float));
      o.n_children = (uint32_t)v[10];
      if (o.n_children) {
        const size_t first = (size_t)v[11];
        if (first + o.n_children > n) throw Napi::Error::New(info.Env(), "render: packed display tree out of range");
        o.children = objs.data() + first;
      }
    }
    stage.n_children = n_root;
    stage.children = objs.data();
    check(info.Env(), swfr_render_display_stage(r_, &stage), "render");
    return info.Env().Undefined();
  }

  Napi::Value ReadImage(const Napi::CallbackInfo &info) {
    const bool premultiplied = info.Length() > 0 && info[0].ToBoolean().Value();
    Napi::ArrayBuffer buf = Napi::ArrayBuffer::New(info.Env(), (size_t)width_ * height_ * 4);
    check(info.Env(), swfr_read_image(r_, 0, static_cast<uint8_t *>(buf.Data()), (size_t)width_ * 4, premultiplied ? 1 : 0), "readImage");
    return Napi::TypedArrayOf<uint8_t>::New(info.Env(), (size_t)width_ * height_ * 4, buf, 0, napi_uint8_clamped_array);
  }

  template <class Writer>
  Napi::Value Encode(const Napi::CallbackInfo &info, Writer write, const char *what) {
    std::vector<uint8_t> px((size_t)width_ * height_ * 4);
    check(info.Env(), swfr_read_image(r_, 0, px.data(), (size_t)width_ * 4, 0), what);  // straight alpha, PNG-export rounding
    uint64_t need = 0;
    check(info.Env(), write(px.data(), width_, height_, (size_t)width_ * 4, nullptr, 0, &need), what);
    Napi::Uint8Array out = Napi::Uint8Array::New(info.Env(), (size_t)need);
    check(info.Env(), write(px.data(), width_, height_, (size_t)width_ * 4, out.Data(), need, &need), what);
    return out;
  }
  Napi::Value ToPng(const Napi::CallbackInfo &info) { return Encode(info, swfr_write_png, "toPng"); }
  Napi::Value ToPam(const Napi::CallbackInfo &info) { return Encode(info, swfr_write_pam, "toPam"); }

  Napi::Value Close(const Napi::CallbackInfo &info) {
    if (r_) swfr_destroy(r_);
    r_ = nullptr;
    return info.Env().Undefined();
  }

  swfr_renderer *r_ = nullptr;
  uint32_t width_ = 0, height_ = 0;
};

Napi::Object InitAll(Napi::Env env, Napi::Object exports) { return Native::Init(env, exports); }

}  // namespace

NODE_API_MODULE(swfr_b200, InitAll)
