// Host-side internal types of libswfr_b200 (shape compiler, stroker, style tables).
// Not part of the ABI; see include/swfr.h for the boundary.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/swfr.h"

namespace swfr {

// ---- device-visible PODs (also used by the kernels) -------------------------------------------------

// One path segment of a static definition: 28 bytes.
struct SegStatic {
  float p[6];           // x0,y0,cx,cy,x1,y1 in twips
  uint32_t path_flags;  // local path index | (is_curve << 31)
};
// One path segment of a morph definition: 52 bytes.
struct SegMorph {
  float s[6];
  float e[6];
  uint32_t path_flags;
};

enum PaintType : uint32_t { PAINT_SOLID = 0, PAINT_LINEAR = 1, PAINT_FOCAL = 2, PAINT_BITMAP = 3 };
enum PaintFlags : uint32_t {
  PF_COLOR_MORPH = 1u,
  PF_OPAQUE_RAMP = 2u,
  PF_SAMPLED = 4u  // stroke outline (overlaps itself): coverage by the non-zero rule per sub-scanline
};

// Per-path paint of a definition ("style table" entry).
struct DefPaint {
  uint32_t type;
  uint32_t spread;
  uint32_t repeating;
  uint32_t flags;
  uint8_t color0[4];
  uint8_t color1[4];
  double matrix[6];  // scale_x, rotate_skew0, rotate_skew1, scale_y, tx, ty  (fill space -> twips)
  double focal;
  int32_t lut;       // index of the gradient's ramp (kRampSize entries) in the ramp store, -1 if none
  uint32_t bitmap_id;
  float bounds[4];   // x_min, y_min, x_max, y_max of the path's control points in twips, both morph states
                     // (x_min > x_max: no segments); the device derives the instance's tile bbox from its corners
};

// Bounds of the control points of every path of a definition's segments -> DefPaint::bounds.
inline void set_paint_bounds(DefPaint *paints, size_t n_paints, const float *seg_points, size_t n_segs, size_t seg_stride_floats,
                             size_t points_per_seg, const uint32_t *path_flags, size_t flags_stride_bytes) {
  for (size_t k = 0; k < n_paints; k++) {
    paints[k].bounds[0] = paints[k].bounds[1] = 1.0f;
    paints[k].bounds[2] = paints[k].bounds[3] = 0.0f;  // empty
  }
  for (size_t i = 0; i < n_segs; i++) {
    const uint32_t pf = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(path_flags) + i * flags_stride_bytes);
    const size_t path = pf & 0x7fffffffu;
    if (path >= n_paints) continue;
    DefPaint &p = paints[path];
    const float *q = seg_points + i * seg_stride_floats;
    for (size_t k = 0; k < points_per_seg; k++) {
      const float x = q[2 * k], y = q[2 * k + 1];
      if (p.bounds[0] > p.bounds[2]) {
        p.bounds[0] = p.bounds[2] = x;
        p.bounds[1] = p.bounds[3] = y;
      } else {
        p.bounds[0] = x < p.bounds[0] ? x : p.bounds[0];
        p.bounds[2] = x > p.bounds[2] ? x : p.bounds[2];
        p.bounds[1] = y < p.bounds[1] ? y : p.bounds[1];
        p.bounds[3] = y > p.bounds[3] ? y : p.bounds[3];
      }
    }
  }
}

// Definition table entry.
struct DefEntry {
  uint32_t seg_first, seg_count;  // into the static or morph segment store
  uint32_t paint_first, path_count;
  uint32_t is_morph;
  uint32_t has_sampled;  // a path of the definition is a stroke outline (PF_SAMPLED)
};

// ---- host-only ---------------------------------------------------------------------------------------

struct Command {  // reference CommandType encoding: LineTo=0, CurveTo=1, MoveTo=2 (path.ts:4-8)
  int type;
  double s[4];  // x|endX, y|endY, controlX, controlY (start state)
  double e[4];  // end state (== s for static shapes)
};

struct CompiledPath {
  std::vector<Command> commands;
  bool has_fill = false, has_line = false;
  swfr_fill_style fill{};   // when has_fill
  swfr_line_style line{};   // when has_line
  std::vector<swfr_color_stop> stops;  // owned copy of the gradient stops
};

// A line path of a morph shape: its outline depends on the ratio (lerped path, lerped width, round caps and joins;
// canvas-renderer.ts:252-266), so it is expanded per draw, not at registration.
struct MorphLine {
  std::vector<Command> commands;  // start + end state
  double w0 = 0, w1 = 0;          // width in twips, start / end
  uint8_t color0[4] = {0, 0, 0, 0}, color1[4] = {0, 0, 0, 0};
};

struct CompiledDef {
  bool is_morph = false;
  std::vector<CompiledPath> paths;       // reference order: per layer fills then lines
  // device form
  std::vector<SegMorph> segs;            // static defs use s only
  std::vector<DefPaint> paints;          // one per emitted device path
  std::vector<std::vector<uint32_t>> luts;  // ramps referenced by paints[].lut (local indices)
  bool has_visible_morph_stroke = false;
  std::vector<MorphLine> morph_lines;    // every line path of a morph shape, in paint order (also invisible ones:
                                         // a zero width keeps the previous one, canvas-renderer.ts:253-256)
};

// Compiles records into ordered style paths and the device segment/paint form.
// Returns a swfr_status; `err` receives a message on failure.
int compile_definition(const swfr_define_shape *tag, bool morph, CompiledDef &out, std::string &err);

// Stroke-to-fill expansion of one path's commands in user space (twips).
// cmds use the same encoding as Command (start state only).  Appends float32-rounded segments.
struct StrokeSeg {
  bool curve;
  float p[6];
};
void stroke_commands(const std::vector<Command> &cmds, double width, bool round_style, std::vector<StrokeSeg> &out);

// Gradient ramp: kRampSize premultiplied RGBA8 entries, looked up without interpolation.
constexpr int kRampSize = 1024;
void build_ramp(const swfr_color_stop *stops, uint32_t n, bool linear_rgb, bool morph_end, std::vector<uint32_t> &out,
                bool *all_opaque);

// image/x-swf-bmp format 3 -> straight RGBA8.  Returns a swfr_status.
int decode_xswfbmp(const uint8_t *data, size_t len, std::vector<uint8_t> &rgba, uint32_t *w, uint32_t *h, std::string &err);
// The host half alone: header check + zlib inflate -> colour table (3 bytes x colors) followed by padded index rows.
int inflate_xswfbmp(const uint8_t *data, size_t len, std::vector<uint8_t> &inflated, uint32_t *w, uint32_t *h, uint32_t *colors,
                    uint32_t *padded, std::string &err);

}  // namespace swfr
