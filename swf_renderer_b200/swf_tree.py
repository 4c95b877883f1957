"""swf-tree 0.8.0 JSON (the corpus' ``ast.json`` files, snake_case serde/kryo form) -> POD descs of swfr.h.

This is the host-side conversion a binding performs before crossing the C ABI; the Rust shim does the same
with ``#[repr(C)]`` mirrors.  Gradient tag names (``linear-gradient`` / ``radial-gradient`` /
``focal-gradient``, ``spread``, ``color_space``) are not exhibited by any reference fixture and follow swf-tree 0.8.0.
"""
from __future__ import annotations

import ctypes as C

from . import capi

_SPREAD = {"pad": capi.SPREAD_PAD, "reflect": capi.SPREAD_REFLECT, "repeat": capi.SPREAD_REPEAT}
_COLOR_SPACE = {"s-rgb": capi.COLOR_SRGB, "linear-rgb": capi.COLOR_LINEAR_RGB}
_FILL = {
    "solid": capi.FILL_SOLID,
    "linear-gradient": capi.FILL_LINEAR_GRADIENT,
    "radial-gradient": capi.FILL_RADIAL_GRADIENT,
    "focal-gradient": capi.FILL_FOCAL_GRADIENT,
    "bitmap": capi.FILL_BITMAP,
}


class Converted:
    """A converted tag plus every ctypes object it points at (kept alive together)."""

    def __init__(self):
        self.keep = []
        self.tag = capi.DefineShape()


def _rgba(c):
    return capi.Rgba8(c["r"], c["g"], c["b"], c["a"])


def _matrix(m):
    return capi.SwfMatrix(
        m["scale_x"], m["scale_y"], m["rotate_skew0"], m["rotate_skew1"], m["translate_x"], m["translate_y"]
    )


def _fill(s, keep) -> capi.FillStyle:
    f = capi.FillStyle()
    t = s["type"]
    if t not in _FILL:
        f.type = 255  # the library reports UnknownFillStyle
        return f
    f.type = _FILL[t]
    if t == "solid":
        f.color = _rgba(s["color"])
        f.morph_color = _rgba(s.get("morph_color", s["color"]))
    elif t == "bitmap":
        f.bitmap_id = s["bitmap_id"]
        f.matrix = _matrix(s["matrix"])
        f.repeating = 1 if s["repeating"] else 0
        f.smoothed = 1 if s["smoothed"] else 0
    else:
        f.matrix = _matrix(s["matrix"])
        g = s["gradient"]
        stops = (capi.ColorStop * max(1, len(g["colors"])))()
        for i, st in enumerate(g["colors"]):
            stops[i].ratio = st["ratio"]
            stops[i].color = _rgba(st["color"])
            stops[i].morph_color = _rgba(st.get("morph_color", st["color"]))
        keep.append(stops)
        f.gradient.spread = _SPREAD[g["spread"]]
        f.gradient.color_space = _COLOR_SPACE[g["color_space"]]
        f.gradient.n_colors = len(g["colors"])
        f.gradient.colors = C.cast(stops, C.POINTER(capi.ColorStop))
        if t == "focal-gradient":
            f.focal_point = s["focal_point"]
    return f


def _styles(st, keep) -> capi.Styles:
    out = capi.Styles()
    fills = (capi.FillStyle * max(1, len(st["fill"])))()
    for i, s in enumerate(st["fill"]):
        fills[i] = _fill(s, keep)
    lines = (capi.LineStyle * max(1, len(st["line"])))()
    for i, s in enumerate(st["line"]):
        lines[i].width = s["width"]
        lines[i].morph_width = s.get("morph_width", s["width"])
        lines[i].fill = _fill(s["fill"], keep)
    keep.extend([fills, lines])
    out.n_fill, out.fill = len(st["fill"]), C.cast(fills, C.POINTER(capi.FillStyle))
    out.n_line, out.line = len(st["line"]), C.cast(lines, C.POINTER(capi.LineStyle))
    return out


def convert_define_shape(tag: dict) -> Converted:
    """``define-shape`` or ``define-morph-shape`` AST dict -> swfr_define_shape."""
    cv = Converted()
    t = cv.tag
    t.id = tag.get("id", 0)
    b = tag["bounds"]
    t.bounds[:] = [b["x_min"], b["x_max"], b["y_min"], b["y_max"]]
    mb = tag.get("morph_bounds", b)
    t.morph_bounds[:] = [mb["x_min"], mb["x_max"], mb["y_min"], mb["y_max"]]
    t.initial_styles = _styles(tag["shape"]["initial_styles"], cv.keep)
    recs = tag["shape"]["records"]
    arr = (capi.ShapeRecord * max(1, len(recs)))()
    for i, r in enumerate(recs):
        o = arr[i]
        if r["type"] == "edge":
            o.type = capi.RECORD_EDGE
            o.delta_x, o.delta_y = r["delta"]["x"], r["delta"]["y"]
            md = r.get("morph_delta", r["delta"])
            o.morph_delta_x, o.morph_delta_y = md["x"], md["y"]
            if r.get("control_delta") is not None:
                o.has_control_delta = 1
                o.control_delta_x, o.control_delta_y = r["control_delta"]["x"], r["control_delta"]["y"]
            if r.get("morph_control_delta") is not None:
                o.has_morph_control_delta = 1
                o.morph_control_delta_x = r["morph_control_delta"]["x"]
                o.morph_control_delta_y = r["morph_control_delta"]["y"]
        elif r["type"] == "style-change":
            o.type = capi.RECORD_STYLE_CHANGE
            if r.get("move_to") is not None:
                o.has_move_to = 1
                o.move_to_x, o.move_to_y = r["move_to"]["x"], r["move_to"]["y"]
            if r.get("morph_move_to") is not None:
                o.has_morph_move_to = 1
                o.morph_move_to_x, o.morph_move_to_y = r["morph_move_to"]["x"], r["morph_move_to"]["y"]
            if r.get("left_fill") is not None:
                o.has_left_fill, o.left_fill = 1, r["left_fill"]
            if r.get("right_fill") is not None:
                o.has_right_fill, o.right_fill = 1, r["right_fill"]
            if r.get("line_style") is not None:
                o.has_line_style, o.line_style = 1, r["line_style"]
            if r.get("new_styles") is not None:
                ns = _styles(r["new_styles"], cv.keep)
                cv.keep.append(ns)
                o.has_new_styles = 1
                o.new_styles = C.pointer(ns)
        else:
            o.type = 255
    cv.keep.append(arr)
    t.n_records = len(recs)
    t.records = C.cast(arr, C.POINTER(capi.ShapeRecord))
    return cv
