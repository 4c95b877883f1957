"""Workloads of BASELINE.json configs 2, 3 and 4, shared by the parity tests
(tests/test_gpu_configs.py) and bench.py - the timed scenes are the tested scenes.

* config 2 `gradients1080`: the `triangle` and `squares` corpus geometries scaled to fill 1920x1080, every fill replaced
  in turn by linear / radial / focal(fp in {-0.75, 0, 0.5}) x spread {pad, reflect, repeat} x colour space {sRGB,
  linear}, 2-, 4- and 15-stop ramps, stop alpha in {255, 128}, gradient rotation {0, 30 deg}: 60 frames, one per case;
* config 3 `morphsweep`: `flat-morph-shapes/homestuck-beta-29` scaled x8 (1072x720), 256 ratios r = 257 k: 256 frames;
* config 4 `textured4k`: the `homestuck-beta-4` quad scaled to 3840x2160 with the corpus bitmap and a seeded 1024x1024
  noise + checker texture, {clipped, repeating} x {smoothed, unsmoothed} x texel:pixel ratio {0.25, 1, 2.58, 8}: 32 frames.

A workload is a `tests/corpus.Scene` (definitions, bitmaps, frames of draw items) - the same object the oracle and
the product are both built from.
"""
import copy
import itertools
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))


def _corpus():
    import corpus

    return corpus


# ---------------------------------------------------------------------------------------------------------
# config 2: gradients at 1920x1080
# ---------------------------------------------------------------------------------------------------------


def gradient_stops(n, alpha, seed):
    rng = np.random.RandomState(seed)
    out = []
    for k in range(n):
        c = rng.randint(0, 256, 3)
        out.append({"ratio": int(round(255 * k / (n - 1))),
                    "color": {"r": int(c[0]), "g": int(c[1]), "b": int(c[2]), "a": int(alpha)}})
    return out


def gradient_fill(kind, focal, spread, space, n_stops, alpha, rot_deg, bounds, seed):
    """Gradient square (-16384..16384) mapped onto the shape's bounding box, rotated by rot_deg."""
    w = bounds["x_max"] - bounds["x_min"]
    h = bounds["y_max"] - bounds["y_min"]
    cx, cy = (bounds["x_max"] + bounds["x_min"]) / 2, (bounds["y_max"] + bounds["y_min"]) / 2
    sx, sy = w / 32768.0, h / 32768.0
    c, s = math.cos(math.radians(rot_deg)), math.sin(math.radians(rot_deg))
    f = {
        "type": kind + "-gradient",
        "matrix": {"scale_x": int(round(sx * c * 65536)), "scale_y": int(round(sy * c * 65536)),
                   "rotate_skew0": int(round(sy * s * 65536)), "rotate_skew1": int(round(-sx * s * 65536)),
                   "translate_x": int(cx), "translate_y": int(cy)},
        "gradient": {"spread": spread, "color_space": space, "colors": gradient_stops(n_stops, alpha, seed)},
    }
    if kind == "focal":
        f["focal_point"] = int(round(focal * 256))
    return f


def fullscreen(tag, W, H):
    """Uniform scale + translate that makes the shape's bounds fill W x H (Matrix2D order)."""
    b = tag["bounds"]
    s = min(W * 20.0 / (b["x_max"] - b["x_min"]), H * 20.0 / (b["y_max"] - b["y_min"]))
    return [s, s, 0.0, 0.0, -b["x_min"] * s, -b["y_min"] * s]


GRAD_KINDS = [("linear", 0.0), ("radial", 0.0), ("focal", -0.75), ("focal", 0.0), ("focal", 0.5)]
GRAD_CASES = []
for _gi, ((_kind, _fp), _spread, _space) in enumerate(
        itertools.product(GRAD_KINDS, ["pad", "reflect", "repeat"], ["s-rgb", "linear-rgb"])):
    # stop count, alpha, rotation and geometry cycle so that the 30 (kind, spread, space) cells cover all of them
    GRAD_CASES.append((["flat-shapes/triangle", "flat-shapes/squares"][_gi % 2], _kind, _fp, _spread, _space,
                       [2, 4, 15][_gi % 3], [255, 128][(_gi // 3) % 2], [0, 30][(_gi // 2) % 2]))
# the complementary choices for the cells the reference can render at all (radial / focal, pad, sRGB) and linear
for _gi, ((_kind, _fp), _n, _alpha, _rot) in enumerate(itertools.product(GRAD_KINDS, [2, 4, 15], [255, 128], [0, 30])):
    if _gi % 2 == 0:
        continue  # half of them: 30 more cases
    GRAD_CASES.append((["flat-shapes/squares", "flat-shapes/triangle"][(_gi // 2) % 2], _kind, _fp, "pad", "s-rgb", _n, _alpha, _rot))


def gradient_tag(geom, kind, fp, spread, space, n_stops, alpha, rot):
    tag = copy.deepcopy(_corpus().load_ast(geom))
    fills = tag["shape"]["initial_styles"]["fill"]
    for i in range(len(fills)):
        fills[i] = gradient_fill(kind, fp, spread, space, n_stops, alpha, rot, tag["bounds"], 1000 + 17 * i + n_stops)
    return tag


def gradient_scene(case, W=1920, H=1080, frame=0, scene=None):
    sc = scene if scene is not None else _corpus().Scene(W, H)
    tag = gradient_tag(*case)
    sc.draw_shape(sc.add_shape(tag), fullscreen(tag, W, H), frame=frame)
    return sc


def gradients1080():
    """60 frames at 1920x1080, one gradient case per frame."""
    sc = _corpus().Scene(1920, 1080)
    for f, case in enumerate(GRAD_CASES):
        gradient_scene(case, 1920, 1080, frame=f, scene=sc)
    return sc


# ---------------------------------------------------------------------------------------------------------
# config 4: bitmap fills at 3840x2160
# ---------------------------------------------------------------------------------------------------------


def noise_checker_1024():
    import synth

    yy, xx = np.mgrid[0:1024, 0:1024]
    n = synth._u(0xC0FFEE, (yy * 1024 + xx).ravel(), 0).reshape(1024, 1024)
    check = (((xx >> 5) + (yy >> 5)) & 1).astype(np.float64)
    img = np.zeros((1024, 1024, 4), dtype=np.uint8)
    img[..., 0] = np.clip(255 * (0.5 * check + 0.5 * n), 0, 255)
    img[..., 1] = np.clip(255 * (xx / 1023.0), 0, 255)
    img[..., 2] = np.clip(255 * (0.6 * (1 - check) + 0.4 * (yy / 1023.0)), 0, 255)
    img[..., 3] = np.where(((xx >> 7) + (yy >> 7)) & 1, 255, 160)
    return img


TEX_CASES = list(itertools.product(["corpus", "noise"], [False, True], [True, False], [0.25, 1.0, 2.58, 8.0]))
NOISE_BITMAP_ID = 7


def textured_scene(case, W=3840, H=2160, frame=0, scene=None):
    """ratio = texels per device pixel (2.58 is the corpus fixture's own minification)."""
    from oracle import decode_bitmap

    corpus = _corpus()
    which, repeating, smoothed, ratio = case
    tag = copy.deepcopy(corpus.load_ast("textured-shapes/homestuck-beta-4"))
    m = fullscreen(tag, W, H)
    sc = scene if scene is not None else corpus.Scene(W, H)
    if which == "corpus":
        bt = corpus.load_bitmap_ast("bitmap/homestuck-beta-3")
        bid = bt["id"]
        if bid not in sc.bitmaps:
            sc.bitmaps[bid] = decode_bitmap.define_bitmap_rgba(bt)
    else:
        bid = NOISE_BITMAP_ID
        if bid not in sc.bitmaps:
            sc.bitmaps[bid] = noise_checker_1024()
    fills = tag["shape"]["initial_styles"]["fill"]
    used = 0
    for i, f in enumerate(fills):
        if f["type"] != "bitmap":
            continue
        # fill matrix maps texels to shape twips; device px per twip = m[0] / 20, so twips per texel:
        tw_per_texel = 20.0 / (m[0] * ratio)
        if i == 1:  # the fill the quad's edges reference (right_fill: 2); fill 0 (id 65535) is never used
            f["bitmap_id"] = bid
            f["matrix"] = {"scale_x": int(round(tw_per_texel * 65536)), "scale_y": int(round(tw_per_texel * 65536)),
                           "rotate_skew0": 0, "rotate_skew1": 0,
                           "translate_x": tag["bounds"]["x_min"] + 400, "translate_y": tag["bounds"]["y_min"] + 300}
            f["repeating"] = repeating
            f["smoothed"] = smoothed
            used += 1
    assert used == 1
    sc.draw_shape(sc.add_shape(tag), m, frame=frame)
    return sc


def textured4k():
    """32 frames at 3840x2160, one bitmap-fill case per frame."""
    sc = _corpus().Scene(3840, 2160)
    for f, case in enumerate(TEX_CASES):
        textured_scene(case, 3840, 2160, frame=f, scene=sc)
    return sc


# ---------------------------------------------------------------------------------------------------------
# config 3: morph-ratio sweep
# ---------------------------------------------------------------------------------------------------------

MORPH_RATIOS = [257 * k for k in range(256)]


def morphsweep(scale=8):
    """256 frames (ratios r = 257 k) of the corpus morph shape at `scale` x its native size (1072x720 at x8)."""
    corpus = _corpus()
    tag = corpus.load_ast(corpus.MORPH_SAMPLE)
    w, h, m = corpus.fixture_canvas(tag)
    ms = [float(scale), float(scale), 0.0, 0.0, m[4] * scale, m[5] * scale]
    sc = corpus.Scene(w * scale, h * scale)
    idx = sc.add_morph(tag)
    for f, r in enumerate(MORPH_RATIOS):
        sc.draw_morph(idx, ms, r, frame=f)
    return sc
