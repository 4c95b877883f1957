"""Oracle restatement of the reference shape compilers (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows, function by function:
  ts/src/lib/shape/decode-swf-shape.ts:22-39     decodeSwfShape
  ts/src/lib/shape/decode-swf-shape.ts:90-149    normalizeStraightSRgba / decodeGradient / decodeFillStyle / decodeLineStyle
  ts/src/lib/shape/decode-swf-shape.ts:179-198   createStyleLayer
  ts/src/lib/shape/decode-swf-shape.ts:203-234   extractContinuous (single forward pass, mutation while iterating)
  ts/src/lib/shape/decode-swf-shape.ts:239-293   segmentsToCommands / layerToPaths
  ts/src/lib/shape/decode-swf-shape.ts:298-448   SwfShapeDecoder
  ts/src/lib/shape/decode-swf-morph-shape.ts     (same, [start, end] pairs; :170-201 matches on start state only;
                                                 :341-346 synthesises a missing control point as delta/2)

Input is the swf-tree 0.8.0 JSON of the fixture corpus (snake_case keys, see tests/*/ast.json).
Output mirrors the TypeScript objects key-for-key so that ``to_golden_json`` reproduces
``JSON.stringify(shape, null, 2) + "\\n"`` byte-exactly (ts/src/test/decode-shape.spec.ts:18-22).
"""
from __future__ import annotations

import json

# ts/src/lib/shape/path.ts:4-8
LINE_TO, CURVE_TO, MOVE_TO = 0, 1, 2
# ts/src/lib/shape/fill-style.ts:5-10
FILL_BITMAP, FILL_FOCAL_GRADIENT, FILL_LINEAR_GRADIENT, FILL_SOLID = 0, 1, 2, 3
# ts/src/lib/shape/morph-fill-style.ts:3-5
MORPH_FILL_SOLID = 0


def _norm_color(c):
    # decode-swf-shape.ts:90-97
    return {"r": c["r"] / 255, "g": c["g"] / 255, "b": c["b"] / 255, "a": c["a"] / 255}


def _matrix(m):
    # swf-tree Matrix as serialised by JSON.stringify: Sfixed16P16 -> {"epsilons": n}
    return {
        "scaleX": {"epsilons": m["scale_x"]},
        "scaleY": {"epsilons": m["scale_y"]},
        "rotateSkew0": {"epsilons": m["rotate_skew0"]},
        "rotateSkew1": {"epsilons": m["rotate_skew1"]},
        "translateX": m["translate_x"],
        "translateY": m["translate_y"],
    }


def _gradient(g):
    # decode-swf-shape.ts:99-105 ; {...swfGradient, colors}
    colors = [{"ratio": s["ratio"] / 0xFF, "color": _norm_color(s["color"])} for s in g["colors"]]
    return {"spread": g["spread"], "colorSpace": g["color_space"], "colors": colors}


def decode_fill_style(s):
    # decode-swf-shape.ts:110-139
    t = s["type"]
    if t == "bitmap":
        return {
            "type": FILL_BITMAP,
            "bitmapId": s["bitmap_id"],
            "matrix": _matrix(s["matrix"]),
            "repeating": s["repeating"],
            "smoothed": s["smoothed"],
        }
    if t == "focal-gradient":
        # Sfixed8P8 focal point: epsilons / 256
        return {
            "type": FILL_FOCAL_GRADIENT,
            "matrix": _matrix(s["matrix"]),
            "gradient": _gradient(s["gradient"]),
            "focalPoint": s["focal_point"] / 256,
        }
    if t == "linear-gradient":
        return {"type": FILL_LINEAR_GRADIENT, "matrix": _matrix(s["matrix"]), "gradient": _gradient(s["gradient"])}
    if t == "radial-gradient":
        return {
            "type": FILL_FOCAL_GRADIENT,
            "matrix": _matrix(s["matrix"]),
            "gradient": _gradient(s["gradient"]),
            "focalPoint": 0,
        }
    if t == "solid":
        return {"type": FILL_SOLID, "color": _norm_color(s["color"])}
    raise ValueError("UnknownFillStyle: %r" % (t,))


def decode_line_style(s):
    # decode-swf-shape.ts:144-149
    return {"width": s["width"], "fill": decode_fill_style(s["fill"])}


def decode_morph_fill_style(s):
    # decode-swf-morph-shape.ts:94-106
    if s["type"] == "solid":
        return {
            "type": MORPH_FILL_SOLID,
            "startColor": _norm_color(s["color"]),
            "endColor": _norm_color(s["morph_color"]),
        }
    raise ValueError("Unknown fill type")


def decode_morph_line_style(s):
    # decode-swf-morph-shape.ts:111-116
    return {"width": [s["width"], s["morph_width"]], "fill": decode_morph_fill_style(s["fill"])}


class _Layer:
    def __init__(self, fills, lines):
        self.fills = [{"style": f, "segments": []} for f in fills]
        self.lines = [{"style": l, "segments": []} for l in lines]


def _extract_continuous(open_set, key):
    """decode-swf-shape.ts:203-234 / decode-swf-morph-shape.ts:170-201.

    ``key(p)`` projects a coordinate to what the reference compares with ``===``
    (the value itself for shapes, element [0] for morph shapes).
    Segments are tuples (sx, sy, cx, cy, ex, ey) with cx is None for straight ones.
    """
    first = open_set.pop(0)
    result = [first]
    start_x, start_y = key(first[0]), key(first[1])
    end_x, end_y = key(first[4]), key(first[5])
    i, n = 0, len(open_set)
    while i < n:
        cur = open_set[i]
        if key(cur[0]) == end_x and key(cur[1]) == end_y:
            del open_set[i]
            i -= 1
            n -= 1
            end_x, end_y = key(cur[4]), key(cur[5])
            result.append(cur)
        elif key(cur[4]) == start_x and key(cur[5]) == start_y:
            del open_set[i]
            i -= 1
            n -= 1
            start_x, start_y = key(cur[0]), key(cur[1])
            result.insert(0, cur)
        i += 1
    return result


def _segments_to_commands(segments, key, wrap):
    # decode-swf-shape.ts:239-273
    open_set = list(segments)
    result = []
    while open_set:
        seq = _extract_continuous(open_set, key)
        result.append({"type": MOVE_TO, "x": wrap(seq[0][0]), "y": wrap(seq[0][1])})
        for s in seq:
            if s[2] is None:
                result.append({"type": LINE_TO, "endX": wrap(s[4]), "endY": wrap(s[5])})
            else:
                result.append(
                    {
                        "type": CURVE_TO,
                        "controlX": wrap(s[2]),
                        "controlY": wrap(s[3]),
                        "endX": wrap(s[4]),
                        "endY": wrap(s[5]),
                    }
                )
    return result


def _layer_to_paths(layer, key, wrap):
    # decode-swf-shape.ts:278-293
    paths = []
    for fs in layer.fills:
        cmds = _segments_to_commands(fs["segments"], key, wrap)
        if cmds:
            paths.append({"commands": cmds, "fill": fs["style"]})
    for ls in layer.lines:
        cmds = _segments_to_commands(ls["segments"], key, wrap)
        if cmds:
            paths.append({"commands": cmds, "line": ls["style"]})
    return paths


def _set_by_id(sets, style_id):
    # decode-swf-shape.ts:410-447: id 0 => none; out of range => throw
    if style_id == 0:
        return None
    idx = style_id - 1
    if idx >= len(sets):
        raise ValueError("Invalid fill ID")
    return sets[idx]


def compile_shape(tag):
    """decodeSwfShape (decode-swf-shape.ts:22-39) on a ``define-shape`` AST dict."""
    shape = tag["shape"]
    layers = []
    left = right = line = None
    x = y = 0

    def new_styles(st):
        nonlocal left, right, line
        layers.append(
            _Layer([decode_fill_style(f) for f in st["fill"]], [decode_line_style(l) for l in st["line"]])
        )
        left = right = line = None

    new_styles(shape["initial_styles"])
    for rec in shape["records"]:
        if rec["type"] == "style-change":
            # decode-swf-shape.ts:337-356: newStyles -> left -> right -> line -> moveTo
            if rec.get("new_styles") is not None:
                new_styles(rec["new_styles"])
            if rec.get("left_fill") is not None:
                left = _set_by_id(layers[-1].fills, rec["left_fill"])
            if rec.get("right_fill") is not None:
                right = _set_by_id(layers[-1].fills, rec["right_fill"])
            if rec.get("line_style") is not None:
                line = _set_by_id(layers[-1].lines, rec["line_style"])
            if rec.get("move_to") is not None:
                x, y = rec["move_to"]["x"], rec["move_to"]["y"]
        elif rec["type"] == "edge":
            # decode-swf-shape.ts:358-390
            ex, ey = x + rec["delta"]["x"], y + rec["delta"]["y"]
            cd = rec.get("control_delta")
            if cd is None:
                cx = cy = None
            else:
                cx, cy = x + cd["x"], y + cd["y"]
            if left is not None:
                left["segments"].append((x, y, cx, cy, ex, ey))
            if right is not None:
                right["segments"].append((ex, ey, cx, cy, x, y))
            if line is not None:
                line["segments"].append((x, y, cx, cy, ex, ey))
            x, y = ex, ey
        else:
            raise ValueError("UnreachableCode")
    paths = []
    for layer in layers:
        paths.extend(_layer_to_paths(layer, lambda v: v, lambda v: v))
    return {"paths": paths}


def compile_morph_shape(tag):
    """decodeSwfMorphShape (decode-swf-morph-shape.ts:21-41) on a ``define-morph-shape`` AST dict."""
    shape = tag["shape"]
    st = shape["initial_styles"]
    layer = _Layer(
        [decode_morph_fill_style(f) for f in st["fill"]], [decode_morph_line_style(l) for l in st["line"]]
    )
    layers = [layer]
    left = right = line = None
    x = (0, 0)
    y = (0, 0)
    for rec in shape["records"]:
        if rec["type"] == "style-change":
            # decode-swf-morph-shape.ts:304-322 (no newStyles handling)
            if rec.get("left_fill") is not None:
                left = _set_by_id(layers[-1].fills, rec["left_fill"])
            if rec.get("right_fill") is not None:
                right = _set_by_id(layers[-1].fills, rec["right_fill"])
            if rec.get("line_style") is not None:
                line = _set_by_id(layers[-1].lines, rec["line_style"])
            if rec.get("move_to") is not None:
                if rec.get("morph_move_to") is None:
                    raise ValueError("Expected morphMoveTo to be defined")
                x = (rec["move_to"]["x"], rec["morph_move_to"]["x"])
                y = (rec["move_to"]["y"], rec["morph_move_to"]["y"])
        elif rec["type"] == "edge":
            # decode-swf-morph-shape.ts:324-364
            d, md = rec["delta"], rec["morph_delta"]
            ex = (x[0] + d["x"], x[1] + md["x"])
            ey = (y[0] + d["y"], y[1] + md["y"])
            cd, mcd = rec.get("control_delta"), rec.get("morph_control_delta")
            if cd is None and mcd is None:
                cx = cy = None
            else:
                if cd is None:
                    cd = {"x": d["x"] / 2, "y": d["y"] / 2}
                if mcd is None:
                    mcd = {"x": md["x"] / 2, "y": md["y"] / 2}
                cx = (x[0] + cd["x"], x[1] + mcd["x"])
                cy = (y[0] + cd["y"], y[1] + mcd["y"])
            if left is not None:
                left["segments"].append((x, y, cx, cy, ex, ey))
            if right is not None:
                right["segments"].append((ex, ey, cx, cy, x, y))
            if line is not None:
                line["segments"].append((x, y, cx, cy, ex, ey))
            x, y = ex, ey
        else:
            raise ValueError("UnreachableCode")
    paths = []
    for lay in layers:
        paths.extend(_layer_to_paths(lay, lambda v: v[0], lambda v: [v[0], v[1]]))
    return {"paths": paths}


# ---------------------------------------------------------------------------------------------
# JSON.stringify(value, null, 2) emulation (numbers printed the JavaScript way)
# ---------------------------------------------------------------------------------------------


def _js_number(v):
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, int):
        return str(v)
    if v == int(v) and abs(v) < 1e21:
        return str(int(v))
    return repr(v)


def _js_dump(v, indent):
    pad = "  " * (indent + 1)
    end = "  " * indent
    if isinstance(v, dict):
        if not v:
            return "{}"
        items = [pad + json.dumps(k) + ": " + _js_dump(x, indent + 1) for k, x in v.items()]
        return "{\n" + ",\n".join(items) + "\n" + end + "}"
    if isinstance(v, (list, tuple)):
        if not v:
            return "[]"
        items = [pad + _js_dump(x, indent + 1) for x in v]
        return "[\n" + ",\n".join(items) + "\n" + end + "]"
    if isinstance(v, str):
        return json.dumps(v)
    if v is None:
        return "null"
    return _js_number(v)


def to_golden_json(shape):
    """``JSON.stringify(shape, null, 2) + "\\n"`` (decode-shape.spec.ts:18)."""
    return _js_dump(shape, 0) + "\n"


# ---------------------------------------------------------------------------------------------
# The Rust crate's decoder (rs/src/decoder/shape_decoder.rs), restated to pin the shape.rs.log goldens
# (rs/src/lib.rs:38-69: `format!("{:#?}\n", decode_shape(&ast.shape))`).  Same layering and single-pass chain
# extraction as the TypeScript compiler, but every segment becomes a straight LineTo (control points are dropped,
# shape_decoder.rs:42-58) and styles are kept as swf-tree values.
# ---------------------------------------------------------------------------------------------


def rust_decode_shape(tag):
    """decode_shape (shape_decoder.rs:18-33): [{"points": [(x, y)], "verbs": [...], "fill": style|None, "line": style|None}]."""
    shape = tag["shape"]

    def new_layer(styles):  # StyleLayerBuilder::new (shape_decoder.rs:181-207)
        return {"fills": [(s, []) for s in styles["fill"]], "lines": [(s, []) for s in styles["line"]], "l": 0, "r": 0, "n": 0}

    layers = []
    top = new_layer(shape["initial_styles"])
    pos = (0, 0)
    for rec in shape["records"]:
        if rec["type"] == "edge":  # apply_edge (shape_decoder.rs:104-109), add_segment (:209-219)
            end = (pos[0] + rec["delta"]["x"], pos[1] + rec["delta"]["y"])
            if top["l"]:
                top["fills"][top["l"] - 1][1].append((pos, end))
            if top["r"]:
                top["fills"][top["r"] - 1][1].append((end, pos))
            if top["n"]:
                top["lines"][top["n"] - 1][1].append((pos, end))
            pos = end
        else:  # apply_style_change (shape_decoder.rs:111-127): new styles, left, right, line, move_to
            if rec.get("new_styles") is not None:
                layers.append(top)
                top = new_layer(rec["new_styles"])
            if rec.get("left_fill") is not None:
                top["l"] = rec["left_fill"]
            if rec.get("right_fill") is not None:
                top["r"] = rec["right_fill"]
            if rec.get("line_style") is not None:
                top["n"] = rec["line_style"]
            if rec.get("move_to") is not None:
                pos = (rec["move_to"]["x"], rec["move_to"]["y"])
    layers.append(top)

    def to_path(segments):  # segments_to_path + extract_continuous (shape_decoder.rs:42-80)
        points, verbs = [], []
        open_set = list(segments)
        while open_set:
            first = open_set.pop(0)
            start, end = first
            chain, remaining = [first], []
            for seg in open_set:
                if seg[0] == end:
                    end = seg[1]
                    chain.append(seg)
                elif seg[1] == start:
                    start = seg[0]
                    chain.insert(0, seg)
                else:
                    remaining.append(seg)
            open_set = remaining
            for k, seg in enumerate(chain):
                if k == 0:
                    points.append(seg[0])
                    verbs.append("MoveTo")
                points.append(seg[1])
                verbs.append("LineTo")
        return points, verbs

    paths = []
    for layer in layers:  # get_shape (shape_decoder.rs:129-165): fills, then lines, empty sets skipped
        for style, segs in layer["fills"]:
            if segs:
                p, v = to_path(segs)
                paths.append({"points": p, "verbs": v, "fill": style, "line": None})
        for style, segs in layer["lines"]:
            if segs:
                p, v = to_path(segs)
                paths.append({"points": p, "verbs": v, "fill": None, "line": style})
    return paths


def _rust_fill(style, ind):
    if style["type"] != "solid":
        raise NotImplementedError("Debug formatting of %s fills (no golden uses them)" % style["type"])
    c = style["color"]
    pad = " " * ind
    return (
        "Solid(\n%s    Solid {\n%s        color: StraightSRgba8 {\n%s            r: %d,\n%s            g: %d,\n"
        "%s            b: %d,\n%s            a: %d,\n%s        },\n%s    },\n%s)"
        % (pad, pad, pad, c["r"], pad, c["g"], pad, c["b"], pad, c["a"], pad, pad, pad)
    )


def rust_debug(paths):
    """`{:#?}` of the Rust `Shape` (derive(Debug) on Shape / StyledPath, lyon's Path, swf-tree's styles)."""
    cap = {"round": "Round", "none": "None", "square": "Square"}
    out = ["Shape {", "    paths: ["]
    for p in paths:
        out += ["        StyledPath {", "            path: Path {", "                points: ["]
        out += ["                    (%s,%s)," % (_rust_f32(x), _rust_f32(y)) for x, y in p["points"]]
        out += ["                ],", "                verbs: ["]
        out += ["                    %s," % v for v in p["verbs"]]
        out += ["                ],", "            },"]
        if p["fill"] is None:
            out.append("            fill: None,")
        else:
            out += ["            fill: Some(", "                " + _rust_fill(p["fill"], 16) + ",", "            ),"]
        if p["line"] is None:
            out.append("            line: None,")
        else:
            ls = p["line"]
            if ls["join"]["type"] != "round":
                raise NotImplementedError("Debug formatting of %s joins (no golden uses them)" % ls["join"]["type"])
            out += ["            line: Some(", "                LineStyle {"]
            out += ["                    width: %d," % ls["width"], "                    start_cap: %s," % cap[ls["start_cap"]],
                    "                    end_cap: %s," % cap[ls["end_cap"]], "                    join: Round,"]
            for k in ("no_h_scale", "no_v_scale", "no_close", "pixel_hinting"):
                out.append("                    %s: %s," % (k, "true" if ls[k] else "false"))
            out += ["                    fill: " + _rust_fill(ls["fill"], 20) + ",", "                },", "            ),"]
        out.append("        },")
    out += ["    ],", "}"]
    return "\n".join(out) + "\n"


def _rust_f32(v):
    """Debug of an f32 holding an integer-valued twips coordinate: one decimal place."""
    f = float(v)
    return "%.1f" % f if f == int(f) else repr(f)
