"""Randomised DefineShape / DefineMorphShape records through the library's host compiler (swfr_compile_debug, no GPU)
against the oracle's restatement of decodeSwfShape / decodeSwfMorphShape: style changes in every field order
combination, newStyles layers, shared left/right fills, open chains, curves, line styles with zero widths - the
record-handling hazards of SURVEY.md appendix A beyond what the five corpus goldens exercise."""
import numpy as np
import pytest

from oracle import compile_shape as cs
from oracle import raster
from test_host_abi import _oracle_commands


def _solid(rng):
    return {"type": "solid", "color": {"r": int(rng.randint(256)), "g": int(rng.randint(256)), "b": int(rng.randint(256)),
                                       "a": int(rng.choice([255, 255, 128, 0]))}}


def _styles(rng, morph):
    def fill():
        f = _solid(rng)
        if morph:
            f["morph_color"] = _solid(rng)["color"]
        return f

    def line():
        ls = {"width": int(rng.choice([0, 20, 40, 100])), "start_cap": "round", "end_cap": "round", "join": {"type": "round"},
              "no_h_scale": False, "no_v_scale": False, "no_close": False, "pixel_hinting": False, "fill": fill()}
        if morph:
            ls["morph_width"] = int(rng.choice([0, 20, 60]))
        return ls

    return {"fill": [fill() for _ in range(rng.randint(1, 4))], "line": [line() for _ in range(rng.randint(0, 3))]}


def _random_tag(seed, morph):
    rng = np.random.RandomState(seed)
    styles = _styles(rng, morph)
    cur = styles
    recs = []
    n = int(rng.randint(3, 40))
    pts = []  # a small pool of points so that chains close and share vertices
    for _ in range(6):
        pts.append((int(rng.randint(-2000, 2000)), int(rng.randint(-2000, 2000))))
    pos = (0, 0)
    mpos = (0, 0)
    have_style = False
    for k in range(n):
        if k == 0 or rng.rand() < 0.3:
            sc = {"type": "style-change"}
            if rng.rand() < 0.7 or k == 0:
                p = pts[rng.randint(len(pts))]
                sc["move_to"] = {"x": p[0], "y": p[1]}
                pos = p
                if morph:
                    q = (p[0] + int(rng.randint(-300, 300)), p[1] + int(rng.randint(-300, 300)))
                    sc["morph_move_to"] = {"x": q[0], "y": q[1]}
                    mpos = q
            if not morph and k > 0 and rng.rand() < 0.15:
                cur = _styles(rng, morph)
                sc["new_styles"] = cur
            if rng.rand() < 0.7 or not have_style:
                sc["left_fill"] = int(rng.randint(0, len(cur["fill"]) + 1))
                have_style = True
            if rng.rand() < 0.5:
                sc["right_fill"] = int(rng.randint(0, len(cur["fill"]) + 1))
            if cur["line"] and rng.rand() < 0.4:
                sc["line_style"] = int(rng.randint(0, len(cur["line"]) + 1))
            recs.append(sc)
        else:
            tgt = pts[rng.randint(len(pts))] if rng.rand() < 0.6 else (int(rng.randint(-2000, 2000)), int(rng.randint(-2000, 2000)))
            d = (tgt[0] - pos[0], tgt[1] - pos[1])
            if d == (0, 0):
                d = (int(rng.randint(1, 500)), int(rng.randint(-500, 500)))
            e = {"type": "edge", "delta": {"x": d[0], "y": d[1]}}
            if rng.rand() < 0.4:
                e["control_delta"] = {"x": d[0] // 2 + int(rng.randint(-200, 200)), "y": d[1] // 2 + int(rng.randint(-200, 200))}
            if morph:
                md = (d[0] + int(rng.randint(-200, 200)), d[1] + int(rng.randint(-200, 200)))
                e["morph_delta"] = {"x": md[0], "y": md[1]}
                if rng.rand() < 0.4:
                    e["morph_control_delta"] = {"x": md[0] // 2 + int(rng.randint(-100, 100)), "y": md[1] // 2 + int(rng.randint(-100, 100))}
                mpos = (mpos[0] + md[0], mpos[1] + md[1])
            pos = (pos[0] + d[0], pos[1] + d[1])
            recs.append(e)
    tag = {"type": "define-morph-shape" if morph else "define-shape", "id": seed & 0xFFFF,
           "bounds": {"x_min": -4000, "x_max": 4000, "y_min": -4000, "y_max": 4000},
           "shape": {"initial_styles": styles, "records": recs}}
    if morph:
        tag["morph_bounds"] = dict(tag["bounds"])
    return tag


@pytest.mark.parametrize("seed", range(60))
def test_random_shapes_compile_like_the_oracle(built_library, seed):
    import swf_renderer_b200 as sw

    tag = _random_tag(seed, False)
    comp = cs.compile_shape(tag)
    cmds, info, segs = sw.compile_tag(tag)
    want = _oracle_commands(comp, False)
    if len(want):
        np.testing.assert_array_equal(cmds, want)
    else:
        assert len(cmds) == 0
    assert [(int(i[0]), bool(i[1]), bool(i[2])) for i in info] == [
        (len(p["commands"]), "fill" in p, "line" in p) for p in comp["paths"]]
    # device segments (implicit close of fills, stroke-to-fill expansion of lines) equal the oracle's
    b = raster._Builder({})
    raster.add_shape_def(b, comp)
    osegs = np.array([[k, path] + list(s6) + list(e6) for (s6, e6, k, path) in b.segs], dtype=np.float64).reshape(-1, 14)
    np.testing.assert_array_equal(segs.reshape(-1, 14), osegs)


@pytest.mark.parametrize("seed", range(100, 140))
def test_random_morph_shapes_compile_like_the_oracle(built_library, seed):
    import swf_renderer_b200 as sw

    tag = _random_tag(seed, True)
    comp = cs.compile_morph_shape(tag)
    cmds, info, segs = sw.compile_tag(tag, morph=True)
    want = _oracle_commands(comp, True)
    if len(want):
        np.testing.assert_array_equal(cmds, want)
    else:
        assert len(cmds) == 0
    assert [(int(i[0]), bool(i[1]), bool(i[2])) for i in info] == [
        (len(p["commands"]), "fill" in p, "line" in p) for p in comp["paths"]]


def _oracle_morph_stroke_segments(tag, ratio):
    """The oracle's expansion of a morph shape's visible strokes at `ratio` (oracle/raster.py: add_morph_shape_item +
    oracle/stroker.py): [(is_curve, line path, x0, y0, cx, cy, x1, y1)]."""
    from oracle import compile_shape as cs
    from oracle import raster

    b = raster._Builder({})
    raster.add_morph_shape_item(b, cs.compile_morph_shape(tag), [1, 1, 0, 0, 0, 0], 0, ratio)
    if not b.defs or b.defs[-1][4]:  # no visible stroke: nothing, or only the fills' (morph) definition, was added
        return [], 0
    first_seg, n_seg, _, n_path, is_morph = b.defs[-1]
    return [(k, lp) + tuple(s6) for (s6, e6, k, lp) in b.segs[first_seg:first_seg + n_seg]], n_path


@pytest.mark.parametrize("ratio", [0.0, 0.25, 0.5, 0.7001953125, 1.0])
def test_device_stroker_generator_matches_the_oracle_stroker(built_library, ratio):
    """the streaming generator the GPU runs per morph draw (csrc/stroke_core.h, compiled for the host here
    through swfr_debug_morph_stroke) yields the oracle stroker's float32 segments bit for bit, in the same order - for
    the corpus morph shape with a visible lerped stroke and for random morph tags with line styles."""
    import copy

    import corpus
    import swf_renderer_b200 as sw
    from swf_renderer_b200.renderer import debug_morph_stroke

    tags = []
    base = copy.deepcopy(corpus.load_ast(corpus.MORPH_SAMPLE))
    for ls in base["shape"]["initial_styles"]["line"]:
        ls["width"], ls["morph_width"] = 60, 140
        ls["fill"] = {"type": "solid", "color": dict(zip("rgba", (0, 0, 0, 255))), "morph_color": dict(zip("rgba", (200, 30, 30, 128)))}
    tags.append(base)
    seed = int(ratio * 1000) * 100 + 5
    while len(tags) < 25:
        t = _random_tag(seed, morph=True)
        seed += 1
        if t["shape"]["initial_styles"]["line"]:
            tags.append(t)
    checked = 0
    for tag in tags:
        try:
            want, want_paths = _oracle_morph_stroke_segments(tag, ratio)
        except NotImplementedError:
            continue
        got, got_paths = debug_morph_stroke(tag, float(np.float32(ratio)))
        assert got_paths == want_paths
        assert len(got) == len(want), (len(got), len(want))
        if len(want):
            np.testing.assert_array_equal(got, np.array(want, dtype=np.float64))
            checked += 1
    assert checked >= 5
    del sw
