"""GPU parity at the BASELINE.json configurations that have no reference fixture:

* config 2 - the `triangle` and `squares` corpus geometries scaled to fill 1920x1080, every fill replaced in turn by
  linear / radial / focal(fp in {-0.75, 0, 0.5}) x spread {pad, reflect, repeat} x colour space {sRGB, linear},
  2-, 4- and 15-stop ramps, stop alpha in {255, 128}, gradient rotation {0, 30 deg};
* config 4 - the `homestuck-beta-4` quad scaled to 3840x2160 with the corpus bitmap and a seeded 1024x1024 noise +
  checker texture, {clipped, repeating} x {smoothed, unsmoothed} x texel:pixel ratio {0.25, 1, 2.58, 8};
* config 3 at throughput size - the morph shape scaled x8 (1072x720), 256 ratios in one batched launch, all frames
  bit-exact against the oracle (native size is covered by test_gpu_parity.py).

The oracle is the checker ("parity unpinned" against the reference for these paints: it has no such fixture and
throws on linear gradients, canvas-renderer.ts:332-333); GPU vs oracle is bit-exact.
"""
import copy
import itertools
import math

import numpy as np
import pytest

import corpus
import synth

pytestmark = pytest.mark.gpu


def _stops(n, alpha, seed):
    rng = np.random.RandomState(seed)
    out = []
    for k in range(n):
        c = rng.randint(0, 256, 3)
        out.append({"ratio": int(round(255 * k / (n - 1))),
                    "color": {"r": int(c[0]), "g": int(c[1]), "b": int(c[2]), "a": int(alpha)}})
    return out


def _gradient_fill(kind, focal, spread, space, n_stops, alpha, rot_deg, bounds, seed):
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
        "gradient": {"spread": spread, "color_space": space, "colors": _stops(n_stops, alpha, seed)},
    }
    if kind == "focal":
        f["focal_point"] = int(round(focal * 256))
    return f


def _fullscreen(tag, W, H):
    """Uniform scale + translate that makes the shape's bounds fill W x H (Matrix2D order)."""
    b = tag["bounds"]
    s = min(W * 20.0 / (b["x_max"] - b["x_min"]), H * 20.0 / (b["y_max"] - b["y_min"]))
    return [s, s, 0.0, 0.0, -b["x_min"] * s, -b["y_min"] * s]


def _render_both(sc):
    ref, info = corpus.render_oracle(sc, want_debug=True)
    r, stages = corpus.make_product(sc)
    r.render(stages[0])
    out = r.get_image(premultiplied=True).data
    np.testing.assert_array_equal(r.debug_tile_counts(0), info["tile_counts"])
    r.close()
    return out, ref


GRAD_KINDS = [("linear", 0.0), ("radial", 0.0), ("focal", -0.75), ("focal", 0.0), ("focal", 0.5)]
GRAD_CASES = []
for gi, ((kind, fp), spread, space) in enumerate(
        itertools.product(GRAD_KINDS, ["pad", "reflect", "repeat"], ["s-rgb", "linear-rgb"])):
    # stop count, alpha, rotation and geometry cycle so that the 30 (kind, spread, space) cells cover all of them
    n_stops = [2, 4, 15][gi % 3]
    alpha = [255, 128][(gi // 3) % 2]
    rot = [0, 30][(gi // 2) % 2]
    geom = ["flat-shapes/triangle", "flat-shapes/squares"][gi % 2]
    GRAD_CASES.append((geom, kind, fp, spread, space, n_stops, alpha, rot))
# the complementary choices for the cells the reference can render at all (radial / focal, pad, sRGB) and linear
for gi, ((kind, fp), n_stops, alpha, rot) in enumerate(
        itertools.product(GRAD_KINDS, [2, 4, 15], [255, 128], [0, 30])):
    if gi % 2 == 0:
        continue  # half of them: 30 more cases
    GRAD_CASES.append((["flat-shapes/squares", "flat-shapes/triangle"][(gi // 2) % 2], kind, fp, "pad", "s-rgb", n_stops, alpha, rot))


@pytest.mark.parametrize("geom,kind,fp,spread,space,n_stops,alpha,rot", GRAD_CASES)
def test_config2_gradients_1080p(built_library, geom, kind, fp, spread, space, n_stops, alpha, rot):
    W, H = 1920, 1080
    tag = copy.deepcopy(corpus.load_ast(geom))
    fills = tag["shape"]["initial_styles"]["fill"]
    seed = hash((geom, kind, fp, spread, space, n_stops, alpha, rot)) & 0x7FFFFFFF
    for i in range(len(fills)):
        fills[i] = _gradient_fill(kind, fp, spread, space, n_stops, alpha, rot, tag["bounds"], 1000 + 17 * i + n_stops)
    del seed
    sc = corpus.Scene(W, H)
    sc.draw_shape(sc.add_shape(tag), _fullscreen(tag, W, H))
    out, ref = _render_both(sc)
    bad = (out != ref).any(axis=2)
    assert not bad.any(), "%d px differ, first %s" % (bad.sum(), np.argwhere(bad)[:4].tolist())
    assert (out[..., 3] > 0).mean() > 0.2  # the shape really fills the frame


def _noise_checker_1024():
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


@pytest.mark.parametrize("which,repeating,smoothed,ratio", TEX_CASES)
def test_config4_textured_4k(built_library, which, repeating, smoothed, ratio):
    """ratio = texels per device pixel (2.58 is the corpus fixture's own minification)."""
    from oracle import decode_bitmap

    W, H = 3840, 2160
    tag = copy.deepcopy(corpus.load_ast("textured-shapes/homestuck-beta-4"))
    m = _fullscreen(tag, W, H)
    sc = corpus.Scene(W, H)
    if which == "corpus":
        bt = corpus.load_bitmap_ast("bitmap/homestuck-beta-3")
        bid = bt["id"]
        sc.bitmaps[bid] = decode_bitmap.define_bitmap_rgba(bt)
    else:
        bid = 7
        sc.bitmaps[bid] = _noise_checker_1024()
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
    sc.draw_shape(sc.add_shape(tag), m)
    out, ref = _render_both(sc)
    bad = (out != ref).any(axis=2)
    assert not bad.any(), "%d px differ, first %s" % (bad.sum(), np.argwhere(bad)[:4].tolist())
    if repeating:
        assert (out[..., 3] > 0).mean() > 0.3


def test_config3_morph_sweep_x8_batched(built_library):
    """256 ratios r = 257 k in ONE batched launch at 8x the native size (1072x720): every frame bit-exact."""
    tag = corpus.load_ast(corpus.MORPH_SAMPLE)
    w, h, m = corpus.fixture_canvas(tag)
    W, H = w * 8, h * 8
    m8 = [8.0, 8.0, 0.0, 0.0, m[4] * 8, m[5] * 8]
    ratios = [257 * k for k in range(256)]
    sc = corpus.Scene(W, H)
    idx = sc.add_morph(tag)
    for f, r in enumerate(ratios):
        sc.draw_morph(idx, m8, r, frame=f)
    from swf_renderer_b200 import capi

    r, stages = corpus.make_product(sc)
    r.set_option(capi.OPT_FRAMES_PER_PASS, 256)  # one set of launches for the whole sweep (and for the edge tap)
    r.render_batch(stages)
    st = r.stats()
    assert st["n_primitives"] == 256
    for f in range(256):
        out = r.get_image(frame=f, premultiplied=True).data
        if f % 8 == 0 or f == 255:  # the oracle takes ~0.1 s per frame at this size: check 33 of them pixel by pixel
            ref, info = corpus.render_oracle(sc, frame=f, want_debug=True)
            assert np.array_equal(out, ref), "ratio %d differs" % ratios[f]
            edges, _ = r.debug_edges(f)
            np.testing.assert_array_equal(edges, info["edges"])
        assert out[..., 3].max() == 255
    r.close()
