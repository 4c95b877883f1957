// Gradient ramps and SWF bitmap decoding (host, at registration time).
//   ramps:   decodeGradient (ts/src/lib/shape/decode-swf-shape.ts:99-105) + Canvas addColorStop semantics
//            (ts/src/lib/renderers/canvas-renderer.ts:327-329)
//   bitmaps: decodeXSwfBmpSync (ts/src/lib/decode-x-swf-bmp.ts:9-41)
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "host_types.h"

namespace swfr {

static double srgb_to_linear(double c) { return c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4); }
static double linear_to_srgb(double c) { return c <= 0.0031308 ? 12.92 * c : 1.055 * std::pow(c, 1.0 / 2.4) - 0.055; }

// kRampSize premultiplied RGBA8 entries; entry k is the gradient at t = (k + 1/2) / kRampSize: straight colour
// interpolated between the stops in double (linear light for linear-RGB gradients), rounded to float, premultiplied
// and quantised in float - the arithmetic the oracle's builder repeats (oracle/raster.py: gradient_lut).
void build_ramp(const swfr_color_stop *stops_in, uint32_t n, bool linear_rgb, bool morph_end, std::vector<uint32_t> &out,
                bool *all_opaque) {
  struct Stop {
    double ratio;
    double c[4];
  };
  std::vector<Stop> stops;
  for (uint32_t i = 0; i < n; i++) {
    const swfr_rgba8 &col = morph_end ? stops_in[i].morph_color : stops_in[i].color;
    Stop s;
    s.ratio = (double)stops_in[i].ratio / 255.0;
    s.c[0] = col.r / 255.0, s.c[1] = col.g / 255.0, s.c[2] = col.b / 255.0, s.c[3] = col.a / 255.0;
    if (linear_rgb)
      for (int k = 0; k < 3; k++) s.c[k] = srgb_to_linear(s.c[k]);
    stops.push_back(s);
  }
  std::stable_sort(stops.begin(), stops.end(), [](const Stop &a, const Stop &b) { return a.ratio < b.ratio; });
  out.assign(kRampSize, 0u);
  bool opaque = n > 0;
  for (const Stop &st : stops)
    if (st.c[3] != 1.0) opaque = false;
  for (int k = 0; k < kRampSize; k++) {
    double t = (k + 0.5) / (double)kRampSize;
    int j = -1;
    for (size_t i = 0; i < stops.size(); i++)
      if (stops[i].ratio <= t) j = (int)i;
    double col[4] = {0, 0, 0, 0};
    if (!stops.empty()) {
      if (j < 0) {
        memcpy(col, stops[0].c, sizeof col);
      } else if (j == (int)stops.size() - 1) {
        memcpy(col, stops[j].c, sizeof col);
      } else {
        const Stop &a = stops[j], &b = stops[j + 1];
        double u = (t - a.ratio) / (b.ratio - a.ratio);
        for (int c = 0; c < 4; c++) col[c] = a.c[c] + (b.c[c] - a.c[c]) * u;
      }
    }
    if (linear_rgb)
      for (int c = 0; c < 3; c++) col[c] = linear_to_srgb(col[c]);
    const float A = (float)col[3];
    auto q8 = [](float v) { return (uint32_t)lrintf(fminf(fmaxf(v * 255.0f, 0.0f), 255.0f)); };
    uint32_t o = q8(A) << 24;
    for (int c = 0; c < 3; c++) o |= q8((float)col[c] * A) << (8 * c);
    out[k] = o;
  }
  if (all_opaque) *all_opaque = opaque;
}

int inflate_xswfbmp(const uint8_t *data, size_t len, std::vector<uint8_t> &src, uint32_t *w, uint32_t *h, uint32_t *colors,
                    uint32_t *padded_out, std::string &err) {
  if (len < 6) {
    err = "x-swf-bmp: truncated header";
    return SWFR_ERR_MALFORMED;
  }
  if (data[0] != 3) {
    err = "UnsupportedXSwfBmpFormatId";
    return SWFR_ERR_UNSUPPORTED_STYLE;
  }
  uint32_t width = data[1] | (data[2] << 8), height = data[3] | (data[4] << 8);
  uint32_t padded = width + ((4 - (width % 4)) % 4);
  uint32_t color_count = (uint32_t)data[5] + 1;
  size_t table = 3 * (size_t)color_count;
  size_t need = table + (size_t)padded * height;
  // The header is untrusted: deflate expands at most ~1032 : 1, so a payload of `len - 6` bytes cannot fill more than
  // that - a larger claim is truncated pixel data, found without allocating (or zero-filling) gigabytes.
  if (need > ((size_t)(len - 6) + 1) * 1032 + 64) {
    err = "x-swf-bmp: pixel data is truncated";
    return SWFR_ERR_MALFORMED;
  }
  src.resize(need);
  uLongf got = (uLongf)need;
  int zr = uncompress(src.data(), &got, data + 6, (uLong)(len - 6));
  if (zr != Z_OK && zr != Z_BUF_ERROR) {
    err = "x-swf-bmp: zlib stream is corrupt";
    return SWFR_ERR_MALFORMED;
  }
  if (got < need) {
    err = "x-swf-bmp: pixel data is truncated";
    return SWFR_ERR_MALFORMED;
  }
  *w = width;
  *h = height;
  *colors = color_count;
  *padded_out = padded;
  return SWFR_OK;
}

int decode_xswfbmp(const uint8_t *data, size_t len, std::vector<uint8_t> &rgba, uint32_t *w, uint32_t *h, std::string &err) {
  std::vector<uint8_t> src;
  uint32_t width = 0, height = 0, color_count = 0, padded = 0;
  int rc = inflate_xswfbmp(data, len, src, &width, &height, &color_count, &padded, err);
  if (rc != SWFR_OK) return rc;
  size_t table = 3 * (size_t)color_count;
  rgba.resize((size_t)width * height * 4);
  for (uint32_t y = 0; y < height; y++) {
    for (uint32_t x = 0; x < width; x++) {
      uint32_t ci = src[table + (size_t)y * padded + x];
      uint8_t *o = &rgba[4 * ((size_t)y * width + x)];
      if (ci < color_count) {
        o[0] = src[3 * ci], o[1] = src[3 * ci + 1], o[2] = src[3 * ci + 2];
      } else {
        o[0] = o[1] = o[2] = 0;  // out-of-range index: opaque black (decode-x-swf-bmp.ts:35-36)
      }
      o[3] = 255;
    }
  }
  *w = width;
  *h = height;
  return SWFR_OK;
}

}  // namespace swfr
