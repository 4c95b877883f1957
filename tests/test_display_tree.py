"""The steps either side of the hot path: the TypeScript display tree flattened into the draw list,
background colour, PAM / PNG emit.  Host parts run without a GPU; the render tests are marked gpu."""
import ctypes as C
import io
import os

import numpy as np
import pytest

import corpus


def _stage_struct(children, keep):
    from swf_renderer_b200 import capi

    def objs(lst):
        arr = (capi.DisplayObject * max(1, len(lst)))()
        keep.append(arr)
        for i, o in enumerate(lst):
            arr[i].type = o["type"]
            arr[i].id = o.get("id", 0)
            if "matrix" in o:
                arr[i].has_matrix = 1
                arr[i].matrix = capi.SwfMatrix(*o["matrix"])
            arr[i].ratio = o.get("ratio", 0.0)
            kids = o.get("children", [])
            if kids:
                k = objs(kids)
                arr[i].n_children = len(kids)
                arr[i].children = C.cast(k, C.POINTER(capi.DisplayObject))
        return arr

    st = capi.DisplayStage()
    st.width, st.height = 100, 100
    root = objs(children)
    st.n_children = len(children)
    st.children = C.cast(root, C.POINTER(capi.DisplayObject))
    return st


def test_flatten_display_tree_order_and_matrices(built_library):
    """Depth-first paint order; a container's matrix applies to its subtree only (save / restore), composed as
    ctx.transform does: CTM' = CTM x M (canvas-renderer.ts:131-145, 179-188)."""
    from swf_renderer_b200 import capi
    from swf_renderer_b200.display import flatten_stage

    S = 65536
    keep = []
    tree = [
        {"type": capi.DISPLAY_SHAPE, "id": 7, "matrix": (S, S, 0, 0, 100, 200)},
        {"type": capi.DISPLAY_CONTAINER, "matrix": (2 * S, 3 * S, 0, 0, 1000, 2000), "children": [
            {"type": capi.DISPLAY_SHAPE, "id": 8},
            {"type": capi.DISPLAY_CONTAINER, "matrix": (0, 0, S, -S, 10, 20), "children": [  # 90 degree rotation
                {"type": capi.DISPLAY_MORPH_SHAPE, "id": 3, "ratio": 0.5, "matrix": (S, S, 0, 0, 5, 7)},
            ]},
            {"type": capi.DISPLAY_SHAPE, "id": 9, "matrix": (S // 2, S // 2, 0, 0, -40, 60)},
        ]},
        {"type": capi.DISPLAY_SHAPE, "id": 10},
    ]
    prims, n = flatten_stage(_stage_struct(tree, keep))
    assert n == 5
    assert [prims[i].id for i in range(n)] == [7, 8, 3, 9, 10]
    assert [prims[i].kind for i in range(n)] == [capi.PRIM_SHAPE, capi.PRIM_SHAPE, capi.PRIM_MORPH_SHAPE, capi.PRIM_SHAPE,
                                                 capi.PRIM_SHAPE]

    def mat(i):
        return [round(float(v), 4) for v in prims[i].matrix]

    # Matrix2D order: scale_x, scale_y, rotate_skew0, rotate_skew1, tx, ty
    assert mat(0) == [1, 1, 0, 0, 100, 200]
    assert mat(1) == [2, 3, 0, 0, 1000, 2000]
    # outer (a=2, d=3, e=1000, f=2000) x rotation (a=0, b=1, c=-1, d=0, e=10, f=20) x translate(5, 7):
    # a = 2*0 + 0*1 = 0, b = 0*0 + 3*1 = 3, c = 2*-1 = -2, d = 0, e = 2*10 + 1000 = 1020, f = 3*20 + 2000 = 2060
    # then translate(5, 7): e' = a*5 + c*7 + e = -14 + 1020 = 1006, f' = b*5 + d*7 + f = 15 + 2060 = 2075
    assert mat(2) == [0, 0, 3, -2, 1006, 2075]
    assert prims[2].flags == capi.PRIM_RATIO_F32 and prims[2].ratio_f == 0.5 and prims[2].ratio == 32768
    assert mat(3) == [1, 1.5, 0, 0, 2 * -40 + 1000, 3 * 60 + 2000]
    assert mat(4) == [1, 1, 0, 0, 0, 0]  # the container's matrix was restored


def test_flatten_rejects_unknown_object_type(built_library):
    from swf_renderer_b200 import capi
    from swf_renderer_b200.display import flatten_stage
    from swf_renderer_b200.renderer import SwfrError

    with pytest.raises(SwfrError):
        flatten_stage(_stage_struct([{"type": 9}], []))


def test_write_pam_matches_reference_golden(built_library):
    """imageDataToPam / write_pam: the reference's own golden (decode-bitmap.spec.ts:31-36) is a PAM of the decoded
    bitmap - decoder + writer together must reproduce the file byte for byte."""
    from swf_renderer_b200.display import image_data_to_pam
    from swf_renderer_b200.renderer import decode_x_swf_bmp

    tag = corpus.load_bitmap_ast("bitmap/homestuck-beta-3")
    rgba = decode_x_swf_bmp(bytes.fromhex(tag["data"]))
    with open(os.path.join(corpus.CORPUS, "bitmap", "homestuck-beta-3.pam"), "rb") as f:
        gold = f.read()
    assert image_data_to_pam(rgba) == gold


def test_write_png_round_trip(built_library):
    from PIL import Image as PILImage

    from swf_renderer_b200.display import write_png

    rng = np.random.RandomState(3)
    img = rng.randint(0, 256, (37, 53, 4)).astype(np.uint8)
    back = np.array(PILImage.open(io.BytesIO(write_png(img))).convert("RGBA"))
    np.testing.assert_array_equal(back, img)
    gold = corpus.load_golden_png("flat-shapes/squares")
    np.testing.assert_array_equal(np.array(PILImage.open(io.BytesIO(write_png(gold))).convert("RGBA")), gold)


@pytest.mark.gpu
def test_display_tree_render_matches_flat_stage_and_oracle(built_library):
    """The TypeScript-shaped API end to end: nested containers, a shape drawn twice from one definition (compiled
    once), a morph shape at float ratio 0.25, PNG export equal to the straight-alpha read-back."""
    from PIL import Image as PILImage

    from swf_renderer_b200 import display as d

    tri = corpus.load_ast("flat-shapes/triangle")
    sq = corpus.load_ast("flat-shapes/squares")
    mo = corpus.load_ast(corpus.MORPH_SAMPLE)
    W, H = 640, 400
    S = 65536
    stage = d.Stage(W, H, [
        d.Shape(tri, d.Matrix(S, S, 0, 0, -380, -820)),
        d.DisplayObjectContainer([
            d.Shape(sq),
            d.DisplayObjectContainer([d.MorphShape(mo, d.Matrix(2 * S, 2 * S, 0, 0, 0, 0), 0.25), d.Shape(sq, d.Matrix(S // 2, S // 2, 0, 0, 0, 0))],
                                     d.Matrix(S, S, S // 4, -S // 4, 2000, 500)),
        ], d.Matrix(S, S, 0, 0, -2000, -1000)),
        d.Shape(tri, d.Matrix(S // 3, S // 3, 0, 0, 4000, 3000)),
    ])
    cr = d.CanvasRenderer(W, H)
    cr.render(stage)
    out = cr.get_image(premultiplied=True).data
    assert len(cr._shape_cache) == 2 and len(cr._morph_cache) == 1
    # the same scene through the oracle, with matrices composed here in float64 and rounded to float32
    from swf_renderer_b200 import capi
    keep = []

    def conv(o):
        e = {}
        if o.matrix is not None:
            m = o.matrix
            e["matrix"] = (m.scale_x, m.scale_y, m.rotate_skew0, m.rotate_skew1, m.translate_x, m.translate_y)
        if isinstance(o, d.DisplayObjectContainer):
            e["type"] = capi.DISPLAY_CONTAINER
            e["children"] = [conv(c) for c in o.children]
        elif isinstance(o, d.MorphShape):
            e["type"], e["ratio"], e["id"] = capi.DISPLAY_MORPH_SHAPE, o.ratio, 0
        else:
            e["type"], e["id"] = capi.DISPLAY_SHAPE, {id(tri): 0, id(sq): 1}[id(o.definition)]
        return e

    prims, n = d.flatten_stage(_stage_struct([conv(c) for c in stage.children], keep))
    sc = corpus.Scene(W, H)
    ti, si, mi = sc.add_shape(tri), sc.add_shape(sq), sc.add_morph(mo)
    for i in range(n):
        m = [float(v) for v in prims[i].matrix]
        if prims[i].kind == capi.PRIM_MORPH_SHAPE:
            sc.draw_morph(mi, m, prims[i].ratio, ratio_f=prims[i].ratio_f)
        else:
            sc.draw_shape([ti, si][prims[i].id], m)
    ref = corpus.render_oracle(sc)
    assert (ref[..., 3] > 0).mean() > 0.1
    np.testing.assert_array_equal(out, ref)
    straight = cr.get_image().data
    np.testing.assert_array_equal(np.array(PILImage.open(io.BytesIO(cr.to_png())).convert("RGBA")), straight)
    assert cr.to_pam().startswith(b"P7\nWIDTH 640\nHEIGHT 400\nDEPTH 4\nMAXVAL 255\nTUPLTYPE RGB_ALPHA\nENDHDR\n")
    cr.close()


@pytest.mark.gpu
def test_background_colour_option(built_library):
    """Default: background ignored, frame starts transparent (canvas-renderer.ts:70-72).  SWFR_OPT_CLEAR_TO_BACKGROUND:
    frame starts from the opaque stage colour (gfx_renderer.rs:292-301)."""
    import swf_renderer_b200 as sw
    from oracle import compile_shape as cs
    from oracle import raster
    from swf_renderer_b200 import capi

    tag = corpus.load_ast("flat-shapes/triangle")
    w, h, m = corpus.fixture_canvas(tag)
    r = sw.HeadlessRenderer(w, h)
    sid = r.register_shape(tag)
    st = sw.Stage([sw.StoredShape(sid, sw.Matrix2D(m))], background_color=(10, 200, 90, 77))
    r.render(st)
    plain = r.get_image(premultiplied=True).data.copy()
    assert (plain[0, 0] == 0).all()
    r.set_option(capi.OPT_CLEAR_TO_BACKGROUND, 1)
    r.render(st)
    out = r.get_image(premultiplied=True).data
    r.close()
    b = raster._Builder({})
    b.add_item(raster.add_shape_def(b, cs.compile_shape(tag)), m)
    ref = raster.render_scene(b.scene(w, h, background=(10, 200, 90)))
    assert tuple(ref[0, 0]) == (10, 200, 90, 255)
    np.testing.assert_array_equal(out, ref)
    assert (out != plain).any()
