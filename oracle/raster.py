"""ctypes front end of the C raster oracle + scene construction (TEST INFRASTRUCTURE, see oracle/__init__.py).

Scene construction restates the host half of the reference draw loop:
  ts/src/lib/renderers/canvas-renderer.ts:269-350   drawPath: beginPath + commands, fill() closes every
                                                    sub-path implicitly; stroke() does not
  ts/src/lib/renderers/canvas-renderer.ts:207-267   drawMorphPath
  ts/src/test/node-canvas-renderer.spec.ts:31-52    fixture canvas size = ceil(bounds/20), matrix = translate(-min)
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

from . import compile_shape as cs
from . import stroker

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile oracle/raster.c with gcc (recipe: oracle/Makefile)."""
    so = os.path.join(_HERE, "liboracle_raster.so")
    src = os.path.join(_HERE, "raster.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle_raster.so"])
    return so


class Segment(C.Structure):
    _fields_ = [("s", C.c_double * 6), ("e", C.c_double * 6), ("is_curve", C.c_int32), ("path", C.c_int32)]


class Paint(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("spread", C.c_int32),
        ("repeating", C.c_int32),
        ("bitmap", C.c_int32),
        ("color0", C.c_uint8 * 4),
        ("color1", C.c_uint8 * 4),
        ("color_is_morph", C.c_int32),
        ("matrix", C.c_double * 6),
        ("focal", C.c_double),
        ("lut", C.POINTER(C.c_uint32)),
        ("sampled", C.c_int32),
    ]


class Def(C.Structure):
    _fields_ = [
        ("first_seg", C.c_int32),
        ("n_seg", C.c_int32),
        ("first_path", C.c_int32),
        ("n_path", C.c_int32),
        ("is_morph", C.c_int32),
    ]


class Item(C.Structure):
    _fields_ = [("def_", C.c_int32), ("m", C.c_float * 6), ("ratio", C.c_uint16), ("use_ratio_f", C.c_uint16),
                ("ratio_f", C.c_float), ("has_cx", C.c_int32), ("cx", C.c_int16 * 8)]


class Bitmap(C.Structure):
    _fields_ = [("w", C.c_int32), ("h", C.c_int32), ("rgba", C.POINTER(C.c_uint8))]


class Scene(C.Structure):
    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("n_items", C.c_int32),
        ("items", C.POINTER(Item)),
        ("defs", C.POINTER(Def)),
        ("segs", C.POINTER(Segment)),
        ("paints", C.POINTER(Paint)),
        ("bitmaps", C.POINTER(Bitmap)),
        ("background", C.c_uint32),
    ]


class Debug(C.Structure):
    _fields_ = [
        ("edges", C.POINTER(C.c_int32)),
        ("edge_path", C.POINTER(C.c_int32)),
        ("edges_cap", C.c_int64),
        ("n_edges", C.c_int64),
        ("tile_counts", C.POINTER(C.c_uint32)),
        ("n_records", C.c_int64),
        ("n_slots_drawn", C.c_int64),
    ]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.swfo_render.argtypes = [C.POINTER(Scene), C.POINTER(C.c_uint8), C.POINTER(Debug)]
        _LIB.swfo_render.restype = C.c_int
        _LIB.swfo_unpremultiply.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        _LIB.swfo_premultiply.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    return _LIB


# ---------------------------------------------------------------------------------------------
# paints
# ---------------------------------------------------------------------------------------------

PAINT_SOLID, PAINT_LINEAR, PAINT_FOCAL, PAINT_BITMAP = 0, 1, 2, 3
SPREAD = {"pad": 0, "reflect": 1, "repeat": 2}


def _srgb_to_linear(c):
    return c / 12.92 if c <= 0.04045 else math.pow((c + 0.055) / 1.055, 2.4)


def _linear_to_srgb(c):
    return 12.92 * c if c <= 0.0031308 else 1.055 * math.pow(c, 1.0 / 2.4) - 0.055


RAMP_SIZE = 1024


def _q8(v):
    """float32 v in 0..1 -> 8 bits: rint(clamp(v * 255)), in float32 like the C side."""
    return int(np.rint(np.minimum(np.maximum(np.float32(v) * np.float32(255.0), np.float32(0.0)), np.float32(255.0))))


def gradient_lut(colors, color_space="s-rgb") -> np.ndarray:
    """RAMP_SIZE premultiplied RGBA8 entries (R | G << 8 | B << 16 | A << 24); entry k is the gradient at
    t = (k + 1/2) / RAMP_SIZE, looked up without interpolation.  ``colors`` = [{"ratio": 0..1, "color": {r,g,b,a in
    0..1}}] as compiled (decode-swf-shape.ts:99-105).  Canvas addColorStop semantics: clamp outside the first/last
    stop, stops kept in insertion order after a stable sort by ratio; at coincident stops the later one wins for
    t >= ratio.  The straight colour is interpolated in double, rounded to float32, premultiplied and quantised in
    float32."""
    stops = sorted(
        [(s["ratio"], [s["color"]["r"], s["color"]["g"], s["color"]["b"], s["color"]["a"]]) for s in colors],
        key=lambda s: s[0],
    )
    linear = color_space == "linear-rgb"
    if linear:
        stops = [(r, [_srgb_to_linear(c[0]), _srgb_to_linear(c[1]), _srgb_to_linear(c[2]), c[3]]) for r, c in stops]
    lut = np.zeros(RAMP_SIZE, dtype=np.uint32)
    for k in range(RAMP_SIZE):
        t = (k + 0.5) / RAMP_SIZE
        j = -1
        for idx, (r, _) in enumerate(stops):
            if r <= t:
                j = idx
        if not stops:
            col = [0.0, 0.0, 0.0, 0.0]
        elif j < 0:
            col = list(stops[0][1])
        elif j == len(stops) - 1:
            col = list(stops[j][1])
        else:
            r0, c0 = stops[j]
            r1, c1 = stops[j + 1]
            u = (t - r0) / (r1 - r0)
            col = [c0[i] + (c1[i] - c0[i]) * u for i in range(4)]
        if linear:
            col = [_linear_to_srgb(col[0]), _linear_to_srgb(col[1]), _linear_to_srgb(col[2]), col[3]]
        a = np.float32(col[3])
        o = _q8(a) << 24
        for c in range(3):
            o |= _q8(np.float32(col[c]) * a) << (8 * c)
        lut[k] = o
    return lut


def _u8(color):
    # normalised colour (c/255 in f64) back to the u8 it came from: exact, see SURVEY appendix A-11
    return [int(round(color[k] * 255)) for k in ("r", "g", "b", "a")]


def _matrix_doubles(m):
    return [
        m["scaleX"]["epsilons"] / 65536.0,
        m["rotateSkew0"]["epsilons"] / 65536.0,
        m["rotateSkew1"]["epsilons"] / 65536.0,
        m["scaleY"]["epsilons"] / 65536.0,
        float(m["translateX"]),
        float(m["translateY"]),
    ]


class _Builder:
    """Accumulates definitions into flat C arrays."""

    def __init__(self, bitmaps):
        self.segs = []  # (s6, e6, is_curve, path)
        self.paints = []
        self.defs = []
        self.items = []
        self.keep = []
        self.bitmap_index = {}
        self.bitmaps = []
        for bid, rgba in (bitmaps or {}).items():
            h, w = rgba.shape[:2]
            pm = np.empty_like(rgba)
            lib().swfo_premultiply(rgba.ctypes.data, pm.ctypes.data, h * w)
            self.keep.append(pm)
            self.bitmap_index[bid] = len(self.bitmaps)
            self.bitmaps.append((w, h, pm))

    def paint_from_fill(self, fill, morph=False):
        p = Paint()
        p.bitmap = -1
        if morph:
            if fill["type"] != cs.MORPH_FILL_SOLID:
                raise NotImplementedError("NotImplementedFillStyle")
            p.type = PAINT_SOLID
            p.color0[:] = _u8(fill["startColor"])
            p.color1[:] = _u8(fill["endColor"])
            p.color_is_morph = 1
            return p
        t = fill["type"]
        if t == cs.FILL_SOLID:
            p.type = PAINT_SOLID
            p.color0[:] = _u8(fill["color"])
            p.color1[:] = _u8(fill["color"])
        elif t == cs.FILL_BITMAP:
            p.type = PAINT_BITMAP
            if fill["bitmapId"] not in self.bitmap_index:
                raise KeyError("BitmapNotFound: %d" % fill["bitmapId"])  # node-canvas-bitmap-service.ts:41-43
            p.bitmap = self.bitmap_index[fill["bitmapId"]]
            p.repeating = 1 if fill["repeating"] else 0
            p.matrix[:] = _matrix_doubles(fill["matrix"])
        elif t in (cs.FILL_FOCAL_GRADIENT, cs.FILL_LINEAR_GRADIENT):
            p.type = PAINT_FOCAL if t == cs.FILL_FOCAL_GRADIENT else PAINT_LINEAR
            p.matrix[:] = _matrix_doubles(fill["matrix"])
            p.focal = float(fill.get("focalPoint", 0))
            p.spread = SPREAD[fill["gradient"]["spread"]]
            lut = gradient_lut(fill["gradient"]["colors"], fill["gradient"]["colorSpace"])
            self.keep.append(lut)
            p.lut = lut.ctypes.data_as(C.POINTER(C.c_uint32))
        else:
            raise NotImplementedError("NotImplementedFillStyle")
        return p

    def add_def(self, paths, is_morph):
        """paths: list of (paint, [(is_curve, s6, e6)])."""
        first_seg, first_path = len(self.segs), len(self.paints)
        for lp, (paint, segs) in enumerate(paths):
            self.paints.append(paint)
            for (is_curve, s6, e6) in segs:
                self.segs.append((s6, e6, is_curve, lp))
        self.defs.append((first_seg, len(self.segs) - first_seg, first_path, len(paths), 1 if is_morph else 0))
        return len(self.defs) - 1

    def add_item(self, def_index, matrix, ratio=0, ratio_f=None, cx=None):
        """cx: a colour transform for this draw - eight integers, red/green/blue/alpha mult (256 = 1.0) then
        red/green/blue/alpha add (include/swfr.h swfr_color_transform)."""
        self.items.append((def_index, matrix, ratio, ratio_f, cx))

    def scene(self, width, height, background=None):
        """background: None = transparent clear (the TypeScript renderer), or (r, g, b) = opaque stage colour."""
        segs = (Segment * max(1, len(self.segs)))()
        for i, (s6, e6, is_curve, path) in enumerate(self.segs):
            segs[i].s[:] = s6
            segs[i].e[:] = e6
            segs[i].is_curve = is_curve
            segs[i].path = path
        paints = (Paint * max(1, len(self.paints)))(*self.paints)
        defs = (Def * max(1, len(self.defs)))()
        for i, d in enumerate(self.defs):
            defs[i].first_seg, defs[i].n_seg, defs[i].first_path, defs[i].n_path, defs[i].is_morph = d
        items = (Item * max(1, len(self.items)))()
        for i, (d, m, r, rf, cx) in enumerate(self.items):
            items[i].def_ = d
            items[i].m[:] = [float(np.float32(v)) for v in m]
            items[i].ratio = r
            if rf is not None:
                items[i].use_ratio_f = 1
                items[i].ratio_f = float(np.float32(rf))
            if cx is not None:
                items[i].has_cx = 1
                items[i].cx[:] = [int(v) for v in cx]
        bitmaps = (Bitmap * max(1, len(self.bitmaps)))()
        for i, (w, h, pm) in enumerate(self.bitmaps):
            bitmaps[i].w, bitmaps[i].h = w, h
            bitmaps[i].rgba = pm.ctypes.data_as(C.POINTER(C.c_uint8))
        sc = Scene()
        sc.width, sc.height, sc.n_items = width, height, len(self.items)
        if background is not None:
            sc.background = int(background[0]) | (int(background[1]) << 8) | (int(background[2]) << 16) | (255 << 24)
        sc.items, sc.defs, sc.segs, sc.paints, sc.bitmaps = items, defs, segs, paints, bitmaps
        self.keep.extend([segs, paints, defs, items, bitmaps])
        return sc


# ---------------------------------------------------------------------------------------------
# compiled paths -> segments
# ---------------------------------------------------------------------------------------------


def _fill_segments(commands):
    """Static fill path: commands -> segments with the implicit close of ctx.fill()."""
    segs = []
    start = cur = None

    def close():
        if start is not None and cur != start:
            segs.append((0, [cur[0], cur[1], cur[0], cur[1], start[0], start[1]]))

    for c in commands:
        if c["type"] == cs.MOVE_TO:
            close()
            start = cur = (float(c["x"]), float(c["y"]))
        elif c["type"] == cs.LINE_TO:
            p = (float(c["endX"]), float(c["endY"]))
            segs.append((0, [cur[0], cur[1], cur[0], cur[1], p[0], p[1]]))
            cur = p
        else:
            p = (float(c["endX"]), float(c["endY"]))
            segs.append((1, [cur[0], cur[1], float(c["controlX"]), float(c["controlY"]), p[0], p[1]]))
            cur = p
    close()
    return [(k, s, list(s)) for k, s in segs]


def _morph_fill_segments(commands):
    segs = []
    start = cur = None

    def pt(xs, ys):
        return ((float(xs[0]), float(ys[0])), (float(xs[1]), float(ys[1])))

    def seg(kind, a, c, b):
        s6 = [a[0][0], a[0][1], c[0][0], c[0][1], b[0][0], b[0][1]]
        e6 = [a[1][0], a[1][1], c[1][0], c[1][1], b[1][0], b[1][1]]
        segs.append((kind, s6, e6))

    def close():
        if start is not None and cur != start:
            seg(0, cur, cur, start)

    for c in commands:
        if c["type"] == cs.MOVE_TO:
            close()
            start = cur = pt(c["x"], c["y"])
        elif c["type"] == cs.LINE_TO:
            p = pt(c["endX"], c["endY"])
            seg(0, cur, cur, p)
            cur = p
        else:
            p = pt(c["endX"], c["endY"])
            seg(1, cur, pt(c["controlX"], c["controlY"]), p)
            cur = p
    close()
    return segs


def _stroke_cmds(commands):
    out = []
    for c in commands:
        if c["type"] == cs.MOVE_TO:
            out.append(("M", (float(c["x"]), float(c["y"]))))
        elif c["type"] == cs.LINE_TO:
            out.append(("L", (float(c["endX"]), float(c["endY"]))))
        else:
            out.append(("Q", (float(c["controlX"]), float(c["controlY"])), (float(c["endX"]), float(c["endY"]))))
    return out


def _lerp(s, e, r):
    return e * r + s * (1 - r)  # canvas-renderer.ts:24-26


def add_shape_def(b: _Builder, compiled):
    """Static shape: one definition, paths in reference order (fills then lines per layer)."""
    paths = []
    width_state = 1.0  # Canvas default lineWidth; a zero width is ignored and the previous one stays
    for path in compiled["paths"]:
        if not path["commands"]:
            continue
        if "fill" in path:
            paths.append((b.paint_from_fill(path["fill"]), _fill_segments(path["commands"])))
        if "line" in path:
            line = path["line"]
            if line["fill"]["type"] != cs.FILL_SOLID:
                raise NotImplementedError("NotImplementedLineStyle")  # canvas-renderer.ts:345-346
            if line["width"] > 0:
                width_state = float(line["width"])
            contours = stroker.stroke_path(_stroke_cmds(path["commands"]), width_state, False)
            segs = [(k, [x0, y0, cx, cy, x1, y1], [x0, y0, cx, cy, x1, y1]) for (k, x0, y0, cx, cy, x1, y1) in stroker.contours_to_segments(contours)]
            paint = b.paint_from_fill(line["fill"])
            paint.sampled = 1  # a stroke outline overlaps itself: non-zero rule per sub-scanline (raster.c: sampled_coverage)
            paths.append((paint, segs))
    return b.add_def(paths, False)


def add_morph_shape_item(b: _Builder, compiled, matrix, ratio_u16, ratio_f=None, cx=None):
    """Morph shape: fills as one morph definition; visible strokes as a transient static definition
    expanded at this ratio (stroke geometry depends on the lerped path and width).  ratio_f (float32, the
    TypeScript renderer's ratio in 0..1) replaces ratio_u16 / 65535 when given."""
    r = ratio_u16 / 65535.0 if ratio_f is None else float(np.float32(ratio_f))
    fills, lines = [], []
    width_state = 1.0
    for path in compiled["paths"]:
        if not path["commands"]:
            continue
        if "fill" in path:
            fills.append((b.paint_from_fill(path["fill"], morph=True), _morph_fill_segments(path["commands"])))
        if "line" in path:
            line = path["line"]
            w = _lerp(float(line["width"][0]), float(line["width"][1]), r)
            if w > 0:
                width_state = w
            cmds = []
            for c in path["commands"]:
                if c["type"] == cs.MOVE_TO:
                    cmds.append(("M", (_lerp(c["x"][0], c["x"][1], r), _lerp(c["y"][0], c["y"][1], r))))
                elif c["type"] == cs.LINE_TO:
                    cmds.append(("L", (_lerp(c["endX"][0], c["endX"][1], r), _lerp(c["endY"][0], c["endY"][1], r))))
                else:
                    cmds.append(
                        (
                            "Q",
                            (_lerp(c["controlX"][0], c["controlX"][1], r), _lerp(c["controlY"][0], c["controlY"][1], r)),
                            (_lerp(c["endX"][0], c["endX"][1], r), _lerp(c["endY"][0], c["endY"][1], r)),
                        )
                    )
            paint = b.paint_from_fill(line["fill"], morph=True)
            paint.sampled = 1
            a = _lerp(paint.color0[3] / 255.0, paint.color1[3] / 255.0, r)
            if a <= 0:
                continue  # invisible stroke: composites nothing
            contours = stroker.stroke_path(cmds, width_state, True)
            segs = [(k, [x0, y0, cx, cy, x1, y1], [x0, y0, cx, cy, x1, y1]) for (k, x0, y0, cx, cy, x1, y1) in stroker.contours_to_segments(contours)]
            lines.append((paint, segs))
    if fills:
        b.add_item(b.add_def(fills, True), matrix, ratio_u16, ratio_f, cx)
    if lines:
        b.add_item(b.add_def(lines, False), matrix, ratio_u16, ratio_f, cx)


# ---------------------------------------------------------------------------------------------
# rendering
# ---------------------------------------------------------------------------------------------


def render_scene(sc: Scene, want_debug=False):
    w, h = sc.width, sc.height
    out = np.zeros((h, w, 4), dtype=np.uint8)
    dbg = Debug()
    keep = []
    if want_debug:
        tiles = ((w + 15) // 16) * ((h + 15) // 16)
        tc = np.zeros(tiles, dtype=np.uint32)
        dbg.tile_counts = tc.ctypes.data_as(C.POINTER(C.c_uint32))
        # first pass to size the edge buffer
        rc = lib().swfo_render(C.byref(sc), out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(dbg))
        assert rc == 0
        n = int(dbg.n_edges)
        edges = np.zeros((max(n, 1), 4), dtype=np.int32)
        epath = np.zeros(max(n, 1), dtype=np.int32)
        dbg.edges = edges.ctypes.data_as(C.POINTER(C.c_int32))
        dbg.edge_path = epath.ctypes.data_as(C.POINTER(C.c_int32))
        dbg.edges_cap = n
        keep = [tc, edges, epath]
    rc = lib().swfo_render(C.byref(sc), out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(dbg))
    if rc != 0:
        raise RuntimeError("swfo_render failed: %d" % rc)
    if want_debug:
        info = {
            "edges": keep[1][: int(dbg.n_edges)],
            "edge_path": keep[2][: int(dbg.n_edges)],
            "tile_counts": keep[0].reshape((h + 15) // 16, (w + 15) // 16),
            "n_records": int(dbg.n_records),
            "n_slots_drawn": int(dbg.n_slots_drawn),
        }
        return out, info
    return out


def unpremultiply(img: np.ndarray) -> np.ndarray:
    out = np.empty_like(img)
    lib().swfo_unpremultiply(np.ascontiguousarray(img).ctypes.data, out.ctypes.data, img.shape[0] * img.shape[1])
    return out


def fixture_canvas(tag):
    """Canvas size and matrix of the reference render test (node-canvas-renderer.spec.ts:31-49, 86-113)."""
    bd = tag["bounds"]
    x_min, x_max, y_min, y_max = bd["x_min"], bd["x_max"], bd["y_min"], bd["y_max"]
    if "morph_bounds" in tag:
        mb = tag["morph_bounds"]
        x_min, x_max = min(x_min, mb["x_min"]), max(x_max, mb["x_max"])
        y_min, y_max = min(y_min, mb["y_min"]), max(y_max, mb["y_max"])
    width = math.ceil((x_max - x_min) / 20)
    height = math.ceil((y_max - y_min) / 20)
    matrix = [1.0, 1.0, 0.0, 0.0, float(-x_min), float(-y_min)]  # Matrix2D order
    return width, height, matrix


def render_shape_fixture(tag, bitmaps=None, want_debug=False, matrix=None, size=None):
    """Render a define-shape AST the way the reference render test does.  Returns premultiplied RGBA8."""
    w, h, m = fixture_canvas(tag)
    if matrix is not None:
        m = matrix
    if size is not None:
        w, h = size
    b = _Builder(bitmaps)
    d = add_shape_def(b, cs.compile_shape(tag))
    b.add_item(d, m)
    return render_scene(b.scene(w, h), want_debug)


def render_morph_fixture(tag, ratio_u16, want_debug=False, matrix=None, size=None):
    w, h, m = fixture_canvas(tag)
    if matrix is not None:
        m = matrix
    if size is not None:
        w, h = size
    b = _Builder(None)
    add_morph_shape_item(b, cs.compile_morph_shape(tag), m, ratio_u16)
    return render_scene(b.scene(w, h), want_debug)
