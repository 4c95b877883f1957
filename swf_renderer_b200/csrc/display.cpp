// Host side of the steps either side of the hot path: the TypeScript renderer's display tree flattened
// into the draw list the kernels consume, and the image writers of the reference.
//   ts/src/lib/display/{stage,display-object,display-object-container,shape,morph-shape}.ts   the tree
//   ts/src/lib/renderers/canvas-renderer.ts:69-94, 114-145, 179-205   renderStage / drawDisplayObject / drawContainer
//                                            (save, applyMatrix, children in order, restore), drawShape, drawMorphShape
//   rs/src/pam.rs:3-34, ts/src/lib/image-data-to-pam.ts:8-30          PAM writer (P7, RGB_ALPHA)
//   ts/src/test/node-canvas-renderer.spec.ts:134-147                  canvas.toBuffer("image/png")
#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/swfr.h"

namespace {

// 2x3 affine in canvas order: x' = a x + c y + e, y' = b x + d y + f
struct Affine {
  double a = 1, b = 0, c = 0, d = 1, e = 0, f = 0;
};

Affine from_swf(const swfr_swf_matrix &m) {  // canvas-renderer.ts:179-188
  Affine r;
  r.a = (double)m.scale_x / 65536.0;
  r.b = (double)m.rotate_skew0 / 65536.0;
  r.c = (double)m.rotate_skew1 / 65536.0;
  r.d = (double)m.scale_y / 65536.0;
  r.e = (double)m.translate_x;
  r.f = (double)m.translate_y;
  return r;
}

// ctx.transform(child) on a context whose matrix is `p`: the child's transform applies first
Affine compose(const Affine &p, const Affine &c) {
  Affine r;
  r.a = p.a * c.a + p.c * c.b;
  r.b = p.b * c.a + p.d * c.b;
  r.c = p.a * c.c + p.c * c.d;
  r.d = p.b * c.c + p.d * c.d;
  r.e = (p.a * c.e + p.c * c.f) + p.e;
  r.f = (p.b * c.e + p.d * c.f) + p.f;
  return r;
}

// SWF colour transforms concatenate down the tree like matrices: the child's applies to the colour first
struct Cx {
  bool on = false;
  int32_t mult[4] = {256, 256, 256, 256}, add[4] = {0, 0, 0, 0};
};

int32_t clamp_i16(int32_t v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

Cx from_swf(const swfr_color_transform &t) {
  Cx r;
  r.on = true;
  const int16_t m[4] = {t.red_mult, t.green_mult, t.blue_mult, t.alpha_mult};
  const int16_t a[4] = {t.red_add, t.green_add, t.blue_add, t.alpha_add};
  for (int i = 0; i < 4; i++) {
    r.mult[i] = m[i];
    r.add[i] = a[i];
  }
  return r;
}

Cx compose(const Cx &p, const Cx &c) {
  if (!p.on) return c;
  Cx r;
  r.on = true;
  for (int i = 0; i < 4; i++) {
    r.mult[i] = clamp_i16((p.mult[i] * c.mult[i]) >> 8);
    r.add[i] = clamp_i16(((p.mult[i] * c.add[i]) >> 8) + p.add[i]);
  }
  return r;
}

int walk(const swfr_display_object *objs, uint32_t n, const Affine &ctm, const Cx &ccx, int depth, swfr_display_primitive *out,
         uint32_t cap, uint32_t &count) {
  if (n && !objs) return SWFR_ERR_INVALID_ARGUMENT;
  if (depth > 256) return SWFR_ERR_INVALID_ARGUMENT;  // a display list is a tree of modest depth; refuse cycles
  for (uint32_t i = 0; i < n; i++) {
    const swfr_display_object &o = objs[i];
    const Affine m = o.has_matrix ? compose(ctm, from_swf(o.matrix)) : ctm;
    const Cx cx = o.has_color_transform ? compose(ccx, from_swf(o.color_transform)) : ccx;
    switch (o.type) {
      case SWFR_DISPLAY_CONTAINER: {
        int rc = walk(o.children, o.n_children, m, cx, depth + 1, out, cap, count);
        if (rc != SWFR_OK) return rc;
        break;
      }
      case SWFR_DISPLAY_SHAPE:
      case SWFR_DISPLAY_MORPH_SHAPE: {
        if (out && count < cap) {
          swfr_display_primitive &p = out[count];
          memset(&p, 0, sizeof p);
          p.kind = o.type == SWFR_DISPLAY_SHAPE ? SWFR_PRIM_SHAPE : SWFR_PRIM_MORPH_SHAPE;
          p.id = o.id;
          // Matrix2D order: scale_x, scale_y, rotate_skew0, rotate_skew1, translate_x, translate_y
          p.matrix[0] = (float)m.a;
          p.matrix[1] = (float)m.d;
          p.matrix[2] = (float)m.b;
          p.matrix[3] = (float)m.c;
          p.matrix[4] = (float)m.e;
          p.matrix[5] = (float)m.f;
          if (o.type == SWFR_DISPLAY_MORPH_SHAPE) {
            p.flags = SWFR_PRIM_RATIO_F32;
            p.ratio_f = o.ratio;
            double q = (double)o.ratio * 65535.0 + 0.5;
            p.ratio = (uint16_t)(q < 0 ? 0 : (q > 65535.0 ? 65535 : (int)q));
          }
          if (cx.on) {
            p.flags |= SWFR_PRIM_COLOR_TRANSFORM;
            p.color_transform = {(int16_t)cx.mult[0], (int16_t)cx.mult[1], (int16_t)cx.mult[2], (int16_t)cx.mult[3],
                                 (int16_t)cx.add[0],  (int16_t)cx.add[1],  (int16_t)cx.add[2],  (int16_t)cx.add[3]};
          }
        }
        count++;
        break;
      }
      default:
        return SWFR_ERR_INVALID_ARGUMENT;  // "UnexpectedDisplayObjectType" (canvas-renderer.ts:91-92)
    }
  }
  return SWFR_OK;
}

void put_be32(std::vector<uint8_t> &v, uint32_t x) {
  v.push_back((uint8_t)(x >> 24));
  v.push_back((uint8_t)(x >> 16));
  v.push_back((uint8_t)(x >> 8));
  v.push_back((uint8_t)x);
}

void png_chunk(std::vector<uint8_t> &out, const char *type, const uint8_t *data, size_t len) {
  put_be32(out, (uint32_t)len);
  size_t at = out.size();
  out.insert(out.end(), type, type + 4);
  if (len) out.insert(out.end(), data, data + len);
  uint32_t crc = (uint32_t)crc32(0L, out.data() + at, (uInt)(len + 4));
  put_be32(out, crc);
}

}  // namespace

extern "C" {

int swfr_flatten_display_stage(const swfr_display_stage *stage, swfr_display_primitive *out, uint32_t cap, uint32_t *n) {
  if (!stage || !n) return SWFR_ERR_INVALID_ARGUMENT;
  uint32_t count = 0;
  int rc;
  try {
    rc = walk(stage->children, stage->n_children, Affine{}, Cx{}, 0, out, cap, count);
  } catch (...) {
    rc = SWFR_ERR_OOM;
  }
  *n = count;
  return rc;
}

int swfr_write_pam(const uint8_t *rgba, uint32_t width, uint32_t height, size_t stride, uint8_t *out, uint64_t cap,
                   uint64_t *n) {
  if (!rgba || !n || stride < (size_t)width * 4) return SWFR_ERR_INVALID_ARGUMENT;
  char hdr[128];
  int hl = snprintf(hdr, sizeof hdr, "P7\nWIDTH %u\nHEIGHT %u\nDEPTH 4\nMAXVAL 255\nTUPLTYPE RGB_ALPHA\nENDHDR\n", width, height);
  uint64_t total = (uint64_t)hl + (uint64_t)width * 4 * height;
  *n = total;
  if (!out || cap < total) return out ? SWFR_ERR_INVALID_ARGUMENT : SWFR_OK;
  memcpy(out, hdr, (size_t)hl);
  for (uint32_t y = 0; y < height; y++) memcpy(out + hl + (size_t)y * width * 4, rgba + (size_t)y * stride, (size_t)width * 4);
  return SWFR_OK;
}

int swfr_write_png(const uint8_t *rgba, uint32_t width, uint32_t height, size_t stride, uint8_t *out, uint64_t cap,
                   uint64_t *n) {
  if (!rgba || !n || width == 0 || height == 0 || stride < (size_t)width * 4) return SWFR_ERR_INVALID_ARGUMENT;
  try {  // no exception may cross the C ABI
  // 8-bit RGBA, non-interlaced, filter type 0 on every row, one IDAT
  std::vector<uint8_t> raw((size_t)height * ((size_t)width * 4 + 1));
  for (uint32_t y = 0; y < height; y++) {
    uint8_t *row = raw.data() + (size_t)y * ((size_t)width * 4 + 1);
    row[0] = 0;
    memcpy(row + 1, rgba + (size_t)y * stride, (size_t)width * 4);
  }
  uLongf zlen = compressBound((uLong)raw.size());
  std::vector<uint8_t> z(zlen);
  if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return SWFR_ERR_OOM;
  std::vector<uint8_t> png;
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  png.insert(png.end(), sig, sig + 8);
  std::vector<uint8_t> ihdr;
  put_be32(ihdr, width);
  put_be32(ihdr, height);
  const uint8_t tail[5] = {8, 6, 0, 0, 0};  // bit depth, colour type RGBA, compression, filter, interlace
  ihdr.insert(ihdr.end(), tail, tail + 5);
  png_chunk(png, "IHDR", ihdr.data(), ihdr.size());
  png_chunk(png, "IDAT", z.data(), (size_t)zlen);
  png_chunk(png, "IEND", nullptr, 0);
  *n = png.size();
  if (!out || cap < png.size()) return out ? SWFR_ERR_INVALID_ARGUMENT : SWFR_OK;
  memcpy(out, png.data(), png.size());
  return SWFR_OK;
  } catch (...) {
    return SWFR_ERR_OOM;
  }
}

}  // extern "C"
