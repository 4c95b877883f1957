"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): flattened edge lists and tile bin counts bit-exact; pixels - this build goes
further than the stated tolerance and requires the premultiplied RGBA8 output to be bit-identical to the
oracle, which itself is pinned on the reference PNGs (tests/test_oracle_golden.py).  The reference goldens are
also checked directly against the GPU output with the north_star tolerance.
"""
import numpy as np
import pytest

import compare
import corpus

pytestmark = pytest.mark.gpu

INTERIOR_TOL = 2  # /255, north_star
PSNR_MIN = 40.0  # dB, north_star


def _assert_same_geometry(r, info, frame=0):
    edges, epath = r.debug_edges(frame)
    np.testing.assert_array_equal(edges, info["edges"])
    np.testing.assert_array_equal(epath, info["edge_path"])
    np.testing.assert_array_equal(r.debug_tile_counts(frame), info["tile_counts"])


@pytest.mark.parametrize("sample,bitmaps", corpus.SHAPE_SAMPLES)
def test_corpus_shapes_match_oracle_and_reference(built_library, sample, bitmaps):
    sc = corpus.corpus_scene(sample, bitmaps)
    ref, info = corpus.render_oracle(sc, want_debug=True)
    r, stages = corpus.make_product(sc)
    r.render(stages[0])
    out = r.get_image(premultiplied=True).data
    _assert_same_geometry(r, info)
    np.testing.assert_array_equal(out, ref)
    # the reference's own golden, north_star tolerance
    gold = corpus.load_golden_png(sample)
    st = compare.stats(out, compare.premultiply_png(gold))
    assert st["psnr"] >= PSNR_MIN and st["interior_max"] <= INTERIOR_TOL, st
    # straight-alpha readback = PNG export rounding
    from oracle import raster

    np.testing.assert_array_equal(r.get_image().data, raster.unpremultiply(ref))
    r.close()


def test_morph_reference_ratios(built_library):
    ratios = [r for r, _ in corpus.MORPH_RATIOS]
    sc = corpus.morph_scene(ratios)
    r, stages = corpus.make_product(sc)
    for f, (ratio, name) in enumerate(corpus.MORPH_RATIOS):
        r.render(stages[f])
        out = r.get_image(premultiplied=True).data
        ref, info = corpus.render_oracle(sc, frame=f, want_debug=True)
        _assert_same_geometry(r, info)
        np.testing.assert_array_equal(out, ref)
        gold = corpus.load_golden_png(corpus.MORPH_SAMPLE, name)
        st = compare.stats(out, compare.premultiply_png(gold))
        assert st["psnr"] >= PSNR_MIN and st["interior_max"] <= 8, st
    r.close()


def test_morph_ratio_sweep_batched(built_library):
    """BASELINE config 3: 256 ratios r = 257 k in one batched call; every frame bit-exact."""
    ratios = [257 * k for k in range(256)]
    sc = corpus.morph_scene(ratios)
    r, stages = corpus.make_product(sc)
    r.set_option(2, 256)  # SWFR_OPT_FRAMES_PER_PASS: the whole sweep in one set of launches
    r.render_batch(stages)
    for f in range(256):
        out = r.get_image(frame=f, premultiplied=True).data
        ref, info = corpus.render_oracle(sc, frame=f, want_debug=True)
        np.testing.assert_array_equal(out, ref, err_msg="ratio %d" % ratios[f])
        edges, _ = r.debug_edges(f)
        np.testing.assert_array_equal(edges, info["edges"])
        np.testing.assert_array_equal(r.debug_tile_counts(f), info["tile_counts"])
    r.close()


def test_empty_stage_and_offscreen(built_library):
    import swf_renderer_b200 as sw

    tag = corpus.load_ast("flat-shapes/triangle")
    r = sw.HeadlessRenderer(100, 60)
    sid = r.register_shape(tag)
    r.render(sw.Stage())
    assert not r.get_image(premultiplied=True).data.any()
    r.render(sw.Stage([sw.StoredShape(sid, sw.Matrix2D.translate(-10_000_000, 0))]))
    assert not r.get_image(premultiplied=True).data.any()
    with pytest.raises(sw.SwfrError) as e:
        r.render(sw.Stage([sw.StoredShape(sid + 7)]))
    assert e.value.status == -2
    r.close()


@pytest.mark.parametrize("dx,dy", [(-3000, -2000), (-9000, 500), (2000, -6000), (4000, 3000)])
def test_viewport_clipping(built_library, dx, dy):
    """Geometry hanging over every viewport edge (backdrop from the left / top must be right)."""
    tag = corpus.load_ast("flat-shapes/triangle")
    w, h, m = corpus.fixture_canvas(tag)
    m = list(m)
    m[4] += dx
    m[5] += dy
    sc = corpus.Scene(301, 173)
    sc.draw_shape(sc.add_shape(tag), m)
    ref, info = corpus.render_oracle(sc, want_debug=True)
    r, stages = corpus.make_product(sc)
    r.render(stages[0])
    _assert_same_geometry(r, info)
    np.testing.assert_array_equal(r.get_image(premultiplied=True).data, ref)
    r.close()


def test_unknown_bitmap_is_an_error(built_library):
    import swf_renderer_b200 as sw

    sc = corpus.corpus_scene("textured-shapes/homestuck-beta-4", None)  # bitmap 3 never registered
    r, stages = corpus.make_product(sc)
    r.render(stages[0])
    with pytest.raises(sw.SwfrError) as e:
        r.sync()
    assert e.value.status == -2  # BitmapNotFound
    r.close()


def _morph_with_visible_strokes(width0=60, width1=140, color0=(0, 0, 0, 255), color1=(200, 30, 30, 128)):
    """The corpus morph shape with its (invisible) line style made visible: lerped width, lerped colour."""
    import copy

    tag = copy.deepcopy(corpus.load_ast(corpus.MORPH_SAMPLE))
    for styles_key in ("initial_styles",):
        for ls in tag["shape"][styles_key]["line"]:
            ls["width"], ls["morph_width"] = width0, width1
            ls["fill"] = {"type": "solid",
                          "color": dict(zip("rgba", color0)), "morph_color": dict(zip("rgba", color1))}
    return tag


@pytest.mark.parametrize("ratio", [0, 1, 16384, 32768, 50000, 65535])
def test_morph_shape_with_visible_strokes(built_library, ratio):
    """morph lines are stroked per draw at the item's ratio (lerped path and width, round caps and
    joins, canvas-renderer.ts:252-266) and painted after the fills - bit-exact against the oracle, which models
    the same expansion (parity unpinned against the reference: its corpus has no visible morph stroke)."""
    tag = _morph_with_visible_strokes()
    w, h, m = corpus.fixture_canvas(tag)
    sc = corpus.Scene(w + 16, h + 16)
    m = [1.0, 1.0, 0.0, 0.0, m[4] + 160.0, m[5] + 160.0]
    sc.draw_morph(sc.add_morph(tag), m, ratio)
    ref, info = corpus.render_oracle(sc, want_debug=True)
    r, stages = corpus.make_product(sc)
    r.render(stages[0])
    out = r.get_image(premultiplied=True).data
    edges, epath = r.debug_edges(0)
    np.testing.assert_array_equal(edges, info["edges"])
    np.testing.assert_array_equal(epath, info["edge_path"])
    np.testing.assert_array_equal(r.debug_tile_counts(0), info["tile_counts"])
    st = r.stats()
    assert st["n_primitives"] == 1 and st["n_path_instances"] >= 2
    r.close()
    bad = (out != ref).any(axis=2)
    assert not bad.any(), "%d px differ" % bad.sum()
    plain = corpus.Scene(sc.width, sc.height)
    plain.draw_morph(plain.add_morph(corpus.load_ast(corpus.MORPH_SAMPLE)), m, ratio)
    assert (corpus.render_oracle(plain) != ref).any()  # the stroke is really visible


def test_morph_strokes_in_a_batched_sweep_and_float_ratio(built_library):
    """Strokes expanded per frame inside one batch (threads of the stage flattener included), and the TypeScript
    renderer's float ratio (0.5 exactly instead of 32768 / 65535)."""
    from swf_renderer_b200 import capi

    tag = _morph_with_visible_strokes(40, 100, (10, 20, 30, 255), (250, 240, 0, 255))
    w, h, m = corpus.fixture_canvas(tag)
    sc = corpus.Scene(w, h)
    idx = sc.add_morph(tag)
    plain = sc.add_morph(corpus.load_ast(corpus.MORPH_SAMPLE))
    ratios = [0, 8192, 30000, 65535]
    for f, rt in enumerate(ratios):
        sc.draw_morph(plain, m, rt, frame=f)
        sc.draw_morph(idx, m, rt, frame=f)
        sc.draw_morph(plain, [0.5, 0.5, 0.0, 0.0, m[4] * 0.5, m[5] * 0.5], 65535 - rt, frame=f)
    sc.draw_morph(idx, m, 0, frame=len(ratios), ratio_f=0.5)
    sc.draw_morph(idx, m, 0, frame=len(ratios) + 1, ratio_f=0.25)
    # several stroked morph shapes in one frame (each draw gets its own transient stroke definition)
    f2 = len(ratios) + 2
    sc.draw_morph(idx, m, 20000, frame=f2)
    sc.draw_morph(idx, [0.6, 0.6, 0.0, 0.0, m[4] * 0.6 + 300.0, m[5] * 0.6 + 200.0], 50000, frame=f2)
    sc.draw_morph(idx, [0.5, 0.5, 0.2, -0.2, m[4] * 0.5 + 900.0, m[5] * 0.5], 65535, frame=f2)
    r, stages = corpus.make_product(sc)
    r.set_option(capi.OPT_FRAMES_PER_PASS, 4)
    r.render_batch(stages)
    for f in range(len(stages)):
        out = r.get_image(frame=f, premultiplied=True).data
        ref = corpus.render_oracle(sc, frame=f)
        assert np.array_equal(out, ref), "frame %d differs in %d px" % (f, (out != ref).any(axis=2).sum())
    r.close()
    # ratio 0.5 as a float is not MorphRatio 32768 (= 0.500008): the two renderings differ somewhere
    a = corpus.Scene(w, h)
    a.draw_morph(a.add_morph(tag), m, 0, ratio_f=0.5)
    b = corpus.Scene(w, h)
    b.draw_morph(b.add_morph(tag), m, 32768)
    assert a.frames != b.frames


def test_xswfbmp_expanded_on_the_device(built_library):
    """Renderer.addBitmap(tag) - the colour table of image/x-swf-bmp is expanded by a kernel: the render
    must equal the one whose bitmap was decoded on the host by the oracle; an index past the table is opaque black
    (decode-x-swf-bmp.ts:35-36); rows are padded to 4 bytes."""
    import zlib

    import swf_renderer_b200 as sw

    sc = corpus.corpus_scene("textured-shapes/homestuck-beta-4", ["bitmap/homestuck-beta-3"])
    ref = corpus.render_oracle(sc)
    r = sw.HeadlessRenderer(sc.width, sc.height)
    r.add_bitmap(corpus.load_bitmap_ast("bitmap/homestuck-beta-3"))
    sid = r.register_shape(sc.shapes[0])
    r.render(sw.Stage([sw.StoredShape(sid, sw.Matrix2D(sc.frames[0][0][2]))]))
    out = r.get_image(premultiplied=True).data
    r.close()
    np.testing.assert_array_equal(out, ref)

    # synthetic 5 x 3 image (rows padded to 8 bytes), 2 colours, some indices out of range
    w, h, colors = 5, 3, [(255, 0, 0), (0, 128, 255)]
    idx = np.array([[0, 1, 2, 1, 0], [1, 1, 0, 255, 7], [0, 0, 1, 1, 1]], dtype=np.uint8)
    rows = np.zeros((h, 8), dtype=np.uint8)
    rows[:, :w] = idx
    raw = bytes(c for col in colors for c in col) + rows.tobytes()
    data = bytes([3, w & 255, w >> 8, h & 255, h >> 8, len(colors) - 1]) + zlib.compress(raw)
    want = np.zeros((h, w, 4), dtype=np.uint8)
    for y in range(h):
        for x in range(w):
            c = colors[idx[y, x]] if idx[y, x] < len(colors) else (0, 0, 0)
            want[y, x] = (*c, 255)
    np.testing.assert_array_equal(sw.decode_x_swf_bmp(data), want)  # host decoder
    # through the device path: fill a rectangle with the bitmap at 1 texel = 1 px, unsmoothed positions at texel centres
    tag = {
        "type": "define-shape", "id": 1, "bounds": {"x_min": 0, "x_max": w * 20, "y_min": 0, "y_max": h * 20},
        "shape": {"initial_styles": {"fill": [{"type": "bitmap", "bitmap_id": 9, "repeating": False, "smoothed": False,
                                               "matrix": {"scale_x": 20 * 65536, "scale_y": 20 * 65536, "rotate_skew0": 0,
                                                          "rotate_skew1": 0, "translate_x": 0, "translate_y": 0}}], "line": []},
                  "records": [{"type": "style-change", "move_to": {"x": 0, "y": 0}, "right_fill": 1},
                              {"type": "edge", "delta": {"x": w * 20, "y": 0}}, {"type": "edge", "delta": {"x": 0, "y": h * 20}},
                              {"type": "edge", "delta": {"x": -w * 20, "y": 0}}, {"type": "edge", "delta": {"x": 0, "y": -h * 20}}]},
    }
    r = sw.HeadlessRenderer(w, h)
    r._check(r._lib.swfr_register_bitmap_xswfbmp(r._h, 9, data, len(data)))
    sid = r.register_shape(tag)
    r.render(sw.Stage([sw.StoredShape(sid)]))
    out = r.get_image(premultiplied=True).data
    r.close()
    np.testing.assert_array_equal(out, want)


def test_gather_frames_between_renderers_in_one_process(built_library):
    """swfr_export_frames / swfr_gather_frames (optional): two renderers of one process shard five frames
    round-robin (frame f -> renderer f mod 2); the first one gathers both stores into global frame order with strided
    asynchronous copies and the result equals rendering all five frames on one renderer."""
    import swf_renderer_b200 as sw
    from swf_renderer_b200.renderer import SwfrError

    sc = corpus.morph_scene([0, 13107, 26214, 39321, 65535])
    r_all, stages = corpus.make_product(sc)
    r_all.render_batch(stages)
    want = np.stack([r_all.get_image(frame=f, premultiplied=True).data.copy() for f in range(5)])
    r_all.close()
    ra, st_a = corpus.make_product(sc)
    rb, st_b = corpus.make_product(sc)
    ra.render_batch([st_a[f] for f in (0, 2, 4)])
    rb.render_batch([st_b[f] for f in (1, 3)])
    out, ms = ra.gather_frames([ra.export_frames(), rb.export_frames()])
    assert out.shape == (5, sc.height, sc.width, 4) and ms >= 0.0
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    # counts that do not fit round-robin sharding are refused
    rb.render_batch([st_b[1]])
    with pytest.raises(SwfrError):
        ra.gather_frames([ra.export_frames(), rb.export_frames()])
    ra.close()
    rb.close()


@pytest.mark.parametrize("ratio_f", [-0.5, 1.4, 2.0])
def test_morph_ratio_outside_unit_interval_extrapolates(built_library, ratio_f):
    """The reference lerps with any number (canvas-renderer.ts:24-26: end * r + start * (1 - r)); a float ratio outside
    [0, 1] extrapolates the geometry past both morph states, and the tile bbox (built from the two states' bounds) grows
    with it instead of cutting the shape off."""
    tag = corpus.load_ast(corpus.MORPH_SAMPLE)
    w, h, m = corpus.fixture_canvas(tag)
    sc = corpus.Scene(3 * w, 3 * h)
    m = [1.0, 1.0, 0.0, 0.0, m[4] + w * 20.0, m[5] + h * 20.0]
    sc.draw_morph(sc.add_morph(tag), m, 0, ratio_f=ratio_f)
    ref, info = corpus.render_oracle(sc, want_debug=True)
    r, stages = corpus.make_product(sc)
    r.render(stages[0])
    out = r.get_image(premultiplied=True).data
    edges, _ = r.debug_edges(0)
    np.testing.assert_array_equal(edges, info["edges"])
    np.testing.assert_array_equal(r.debug_tile_counts(0), info["tile_counts"])
    r.close()
    assert np.array_equal(out, ref) and out[..., 3].any()


def test_device_stroker_outgrown_room_is_laid_out_again(built_library):
    """the outline of a morph stroke is generated on the device into room reserved at registration; with
    SWFR_OPT_DEBUG_TINY_ARENA the room is far too small, k_stroke reports it, the host lays the batch out again with
    exact counts and the render still equals the oracle - also for several draws and frames in one batch."""
    from swf_renderer_b200 import capi

    tag = _morph_with_visible_strokes(40, 100, (10, 20, 30, 255), (250, 240, 0, 255))
    w, h, m = corpus.fixture_canvas(tag)
    sc = corpus.Scene(w, h)
    idx = sc.add_morph(tag)
    for f, rt in enumerate([0, 21845, 43690, 65535]):
        sc.draw_morph(idx, m, rt, frame=f)
        sc.draw_morph(idx, [0.5, 0.5, 0.0, 0.0, m[4] * 0.5 + 200.0, m[5] * 0.5 + 100.0], 65535 - rt, frame=f)
    r, stages = corpus.make_product(sc)
    r.set_option(capi.OPT_FRAMES_PER_PASS, 2)
    r.set_option(capi.OPT_DEBUG_TINY_ARENA, 1)
    r.render_batch(stages)
    assert r.stats()["retries"] >= 1
    for f in range(4):
        ref = corpus.render_oracle(sc, frame=f)
        assert np.array_equal(r.get_image(frame=f, premultiplied=True).data, ref), "frame %d" % f
    r.set_option(capi.OPT_DEBUG_TINY_ARENA, 0)
    r.render_batch(stages)  # the estimate was raised: no retry now
    assert r.stats()["retries"] == 0
    for f in range(4):
        assert np.array_equal(r.get_image(frame=f, premultiplied=True).data, corpus.render_oracle(sc, frame=f))
    r.close()
