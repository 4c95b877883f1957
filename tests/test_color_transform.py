"""Colour transforms (swf-tree ColorTransformWithAlpha) on draws: not an input of the reference renderer (its display
list carries a matrix and a ratio only), so the semantics are the oracle's own ("parity unpinned", include/swfr.h
swfr_color_transform).  Host parts run without a GPU; the product-vs-oracle tests are marked gpu."""
import copy
import ctypes as C

import numpy as np
import pytest

import corpus

IDENTITY = (256, 256, 256, 256, 0, 0, 0, 0)


def cx_channel(c, mult, add):
    return min(255, max(0, ((c * mult) >> 8) + add))


def cx_premul_px(px, cx):
    """Independent restatement for one premultiplied RGBA pixel."""
    r, g, b, a = [int(v) for v in px]
    a2 = cx_channel(a, cx[3], cx[7])
    out = []
    for k, c in enumerate((r, g, b)):
        s = min(255, (c * 255 + a // 2) // a) if a else 0
        out.append((cx_channel(s, cx[k], cx[4 + k]) * a2 + 127) // 255)
    return out + [a2]


def _solid_square(color):
    tag = copy.deepcopy(corpus.load_ast("flat-shapes/squares"))
    for f in tag["shape"]["initial_styles"]["fill"]:
        f["color"] = {"r": color[0], "g": color[1], "b": color[2], "a": color[3]}
    return tag


def _one_draw(tag, cx, W=160, H=120):
    import workloads

    sc = corpus.Scene(W, H)
    sc.draw_shape(sc.add_shape(tag), workloads.fullscreen(tag, W, H), cx=cx)
    return sc


# ---------------------------------------------------------------------------------------------------------
# oracle semantics (CPU)
# ---------------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("color,cx,expect", [
    # mult 0.5 on red: (255 * 128) >> 8 = 127; opaque stays opaque
    ((255, 255, 255, 255), (128, 256, 256, 256, 0, 0, 0, 0), (127, 255, 255, 255)),
    # add saturates, negative add clamps at 0
    ((200, 100, 50, 255), (256, 256, 256, 256, 100, -200, 0, 0), (255, 0, 50, 255)),
    # alpha mult 0.5: alpha 127, colours premultiplied by it (Cairo's 16-bit path: (c / 255 * a) * 65535 + 0.5 >> 8)
    ((255, 0, 0, 255), (256, 256, 256, 128, 0, 0, 0, 0), (127, 0, 0, 127)),
    # negative mult floors: (100 * -256) >> 8 = -100, + 255 = 155
    ((100, 100, 100, 255), (-256, 256, 256, 256, 255, 0, 0, 0), (155, 100, 100, 255)),
])
def test_oracle_solid_fill_transform(color, cx, expect):
    img = corpus.render_oracle(_one_draw(_solid_square(color), cx))
    inside = img[img[..., 3] == expect[3]]
    assert len(inside) > 1000
    vals, counts = np.unique(inside.reshape(-1, 4), axis=0, return_counts=True)
    assert tuple(int(v) for v in vals[counts.argmax()]) == expect


def test_oracle_identity_transform_is_no_transform():
    import workloads

    tag = workloads.gradient_tag("flat-shapes/squares", "radial", 0.0, "pad", "s-rgb", 4, 128, 30)
    plain = corpus.render_oracle(_one_draw(tag, None))
    np.testing.assert_array_equal(corpus.render_oracle(_one_draw(tag, IDENTITY)), plain)
    assert (corpus.render_oracle(_one_draw(tag, (256, 256, 256, 256, 0, 0, 1, 0))) != plain).any()


def test_oracle_gradient_pixels_go_through_unpremultiply_transform_premultiply():
    """Where the gradient covers a pixel completely (mask 255 onto a transparent canvas) the result is the transform of
    the untransformed render's pixel."""
    import workloads

    tag = workloads.gradient_tag("flat-shapes/squares", "linear", 0.0, "pad", "s-rgb", 4, 128, 0)
    cx = (300, 200, 256, 180, -20, 10, 40, 30)
    plain = corpus.render_oracle(_one_draw(tag, None))
    moved = corpus.render_oracle(_one_draw(tag, cx))
    # the squares fixture draws disjoint squares: interior pixels have the gradient's own alpha (128 -> premultiplied)
    ys, xs = np.nonzero(plain[..., 3] == 128)
    assert len(ys) > 1000
    for y, x in list(zip(ys, xs))[::97]:
        assert list(moved[y, x]) == cx_premul_px(plain[y, x], cx), (y, x)


def test_flatten_concatenates_color_transforms_down_the_tree(built_library):
    from swf_renderer_b200 import capi
    from swf_renderer_b200.display import flatten_stage

    def ct(*v):
        return capi.ColorTransform(*v)

    keep = []
    arr_leaf = (capi.DisplayObject * 2)()
    arr_leaf[0].type, arr_leaf[0].id = capi.DISPLAY_SHAPE, 1
    arr_leaf[0].has_color_transform = 1
    arr_leaf[0].color_transform = ct(128, 256, 512, 256, 10, 0, -5, 0)
    arr_leaf[1].type, arr_leaf[1].id = capi.DISPLAY_SHAPE, 2  # inherits the container's
    arr_root = (capi.DisplayObject * 2)()
    arr_root[0].type = capi.DISPLAY_CONTAINER
    arr_root[0].has_color_transform = 1
    arr_root[0].color_transform = ct(128, 128, 128, 200, 100, -100, 0, 7)
    arr_root[0].n_children = 2
    arr_root[0].children = C.cast(arr_leaf, C.POINTER(capi.DisplayObject))
    arr_root[1].type, arr_root[1].id = capi.DISPLAY_SHAPE, 3  # outside: none
    st = capi.DisplayStage()
    st.width, st.height, st.n_children = 10, 10, 2
    st.children = C.cast(arr_root, C.POINTER(capi.DisplayObject))
    keep += [arr_leaf, arr_root]
    prims, n = flatten_stage(st)
    assert n == 3

    def cx(i):
        t = prims[i].color_transform
        return (t.red_mult, t.green_mult, t.blue_mult, t.alpha_mult, t.red_add, t.green_add, t.blue_add, t.alpha_add)

    assert prims[0].flags & capi.PRIM_COLOR_TRANSFORM and prims[1].flags & capi.PRIM_COLOR_TRANSFORM
    assert not prims[2].flags & capi.PRIM_COLOR_TRANSFORM
    # child first: mult = (Pm * Cm) >> 8, add = ((Pm * Ca) >> 8) + Pa
    assert cx(0) == (64, 128, 256, 200, 105, -100, -3, 7)  # (128 * -5) >> 8 = -3 (floor)
    assert cx(1) == (128, 128, 128, 200, 100, -100, 0, 7)


# ---------------------------------------------------------------------------------------------------------
# product vs oracle (GPU)
# ---------------------------------------------------------------------------------------------------------

CXS = [
    (128, 256, 256, 256, 0, 0, 0, 0),
    (256, 256, 256, 128, 0, 0, 0, 0),        # opaque paints become translucent: no occlusion culling by them
    (300, 200, 256, 180, -20, 10, 40, 30),
    (-256, -256, -256, 256, 255, 255, 255, 0),  # inversion
    (256, 256, 256, 0, 0, 0, 0, 255),        # alpha forced to 255: translucent paints become opaque
    (256, 256, 256, 256, 0, 0, 0, -255),     # everything transparent
    (32767, -32768, 256, 256, -32768, 32767, 0, 0),
]


def _render_product(sc):
    r, stages = corpus.make_product(sc)
    r.render_batch(stages)
    out = [r.get_image(frame=f, premultiplied=True).data.copy() for f in range(len(stages))]
    r.close()
    return out


def _check(sc):
    out = _render_product(sc)
    for f in range(len(sc.frames)):
        np.testing.assert_array_equal(out[f], corpus.render_oracle(sc, frame=f), err_msg=f"frame {f}")
    return out


@pytest.mark.gpu
def test_solid_fills_and_strokes_with_color_transforms(built_library):
    sc = corpus.Scene(320, 240)
    import workloads

    tags = [corpus.load_ast(s) for s in ("flat-shapes/homestuck-beta-1", "flat-shapes/squares", "flat-shapes/triangle")]
    ids = [sc.add_shape(t) for t in tags]
    for f, cx in enumerate(CXS):
        for k, (i, t) in enumerate(zip(ids, tags)):
            m = workloads.fullscreen(t, 320, 240)
            # overlapping draws, some transformed, some not, one with the identity
            sc.draw_shape(i, m, frame=f, cx=[cx, None, IDENTITY][(f + k) % 3] if k else cx)
    out = _check(sc)
    assert (out[0] != out[2]).any()


@pytest.mark.gpu
def test_gradients_with_color_transforms(built_library):
    import workloads

    W, H = 384, 256
    sc = corpus.Scene(W, H)
    cases = [workloads.GRAD_CASES[i] for i in (0, 7, 14, 21, 28, 35, 44)]
    for f, (case, cx) in enumerate(zip(cases, CXS)):
        tag = workloads.gradient_tag(*case)
        m = workloads.fullscreen(tag, W, H)
        i = sc.add_shape(tag)
        sc.draw_shape(i, m, frame=f)                                       # plain underneath
        sc.draw_shape(i, [m[0] * 0.8, m[1] * 0.8, 0.1, -0.1, m[4] + 300, m[5] + 200], frame=f, cx=cx)
    _check(sc)


@pytest.mark.gpu
def test_opaque_gradient_under_alpha_transform_does_not_cull(built_library):
    """An opaque gradient hides what is below it (occlusion culling); with an alpha transform it no longer does, and a
    translucent one whose alpha is forced to 255 does."""
    import workloads

    W, H = 512, 384
    opaque = workloads.gradient_tag("flat-shapes/squares", "linear", 0.0, "pad", "s-rgb", 4, 255, 0)
    translucent = workloads.gradient_tag("flat-shapes/squares", "radial", 0.0, "pad", "s-rgb", 4, 128, 30)
    under = corpus.load_ast("flat-shapes/triangle")
    sc = corpus.Scene(W, H)
    iu, io, it = sc.add_shape(under), sc.add_shape(opaque), sc.add_shape(translucent)
    for f, (top, cx) in enumerate([(io, None), (io, (256, 256, 256, 128, 0, 0, 0, 0)), (it, None),
                                   (it, (256, 256, 256, 0, 0, 0, 0, 255))]):
        sc.draw_shape(iu, workloads.fullscreen(under, W, H), frame=f)
        t = opaque if top == io else translucent
        m = workloads.fullscreen(t, W, H)
        sc.draw_shape(top, [m[0] * 4, m[1] * 4, 0, 0, m[4] * 4 - 600, m[5] * 4 - 600], frame=f, cx=cx)  # covers whole tiles
    out = _check(sc)
    assert (out[0] != out[1]).any() and (out[2] != out[3]).any()


@pytest.mark.gpu
def test_bitmap_fills_with_color_transforms(built_library):
    import workloads

    W, H = 480, 270
    sc = corpus.Scene(W, H)
    cases = [("corpus", False, True, 1.0), ("noise", True, True, 2.58), ("noise", True, False, 0.25), ("corpus", True, True, 8.0)]
    for f, case in enumerate(cases):
        workloads.textured_scene(case, W, H, frame=f, scene=sc)
        kind, idx, m, ratio, _ = sc.frames[f][-1]
        sc.frames[f][-1] = (kind, idx, m, ratio, CXS[(f + 2) % len(CXS)])
    # a non-repeating bitmap whose transparent outside gets alpha added: the transform applies to every evaluated pixel
    workloads.textured_scene(("corpus", False, True, 1.0), W, H, frame=len(cases), scene=sc)
    kind, idx, m, ratio, _ = sc.frames[len(cases)][-1]
    sc.frames[len(cases)][-1] = (kind, idx, m, ratio, (256, 256, 256, 256, 0, 60, 0, 90))
    _check(sc)


@pytest.mark.gpu
def test_morph_shapes_with_color_transforms(built_library):
    """Morph fills (lerped solid colours) and device-stroked morph lines under a transform, mixed with plain draws; one
    frame without any transform in the same batch (its pass variant is chosen per pass)."""
    from test_gpu_parity import _morph_with_visible_strokes

    tag = _morph_with_visible_strokes(40, 100, (10, 20, 30, 255), (250, 240, 0, 128))
    w, h, m = corpus.fixture_canvas(tag)
    sc = corpus.Scene(w, h)
    idx = sc.add_morph(tag)
    plain = sc.add_morph(corpus.load_ast(corpus.MORPH_SAMPLE))
    for f, cx in enumerate(CXS[:5]):
        sc.draw_morph(plain, m, 10000 * f, frame=f)
        sc.draw_morph(idx, m, 13107 * f, frame=f, cx=cx)
        sc.draw_morph(idx, [m[0] * 0.5, m[1] * 0.5, 0, 0, m[4] * 0.5, m[5] * 0.5], 0, frame=f, ratio_f=0.37, cx=CXS[(f + 3) % len(CXS)])
    sc.draw_morph(idx, m, 30000, frame=5)
    _check(sc)


@pytest.mark.gpu
def test_display_tree_color_transform_end_to_end(built_library):
    """The TypeScript-shaped API: a container's transform applies to its subtree, concatenated with the child's."""
    from swf_renderer_b200 import display as d

    tri = corpus.load_ast("flat-shapes/triangle")
    sq = corpus.load_ast("flat-shapes/squares")
    W, H, S = 400, 300, 65536
    outer = d.ColorTransform(128, 256, 256, 200, 40, 0, 0, 0)
    inner = d.ColorTransform(256, 128, 256, 256, 0, 0, 90, 0)
    stage = d.Stage(W, H, [
        d.Shape(sq),
        d.DisplayObjectContainer([d.Shape(tri, d.Matrix(S, S, 0, 0, -380, -820)), d.Shape(sq, d.Matrix(S // 2, S // 2, 0, 0, 900, 700), inner)],
                                 d.Matrix(S, S, 0, 0, 200, 100), outer),
    ])
    cr = d.CanvasRenderer(W, H)
    cr.render(stage)
    out = cr.get_image(premultiplied=True).data
    cr.close()
    both = ((128 * 256) >> 8, (256 * 128) >> 8, 256, (200 * 256) >> 8, 40, 0, ((256 * 90) >> 8), 0)
    sc = corpus.Scene(W, H)
    ti, si = sc.add_shape(tri), sc.add_shape(sq)
    sc.draw_shape(si, [1, 1, 0, 0, 0, 0])
    sc.draw_shape(ti, [1, 1, 0, 0, 200 - 380, 100 - 820], cx=(128, 256, 256, 200, 40, 0, 0, 0))
    sc.draw_shape(si, [0.5, 0.5, 0, 0, 200 + 900, 100 + 700], cx=both)
    np.testing.assert_array_equal(out, corpus.render_oracle(sc))
