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
