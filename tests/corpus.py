"""Access to the committed copy of the reference's fixture corpus (tests/golden/corpus, copied from the
reference repository's tests/ directory: inputs ast.json, goldens shape.ts.json / *.png / *.pam).

Also builds the same scene twice - once for the oracle, once for the product - so parity tests compare like
with like.  Canvas size and matrix follow ts/src/test/node-canvas-renderer.spec.ts:31-52, 86-113.
"""
import json
import math
import os

import numpy as np
from PIL import Image as PILImage

HERE = os.path.dirname(os.path.abspath(__file__))
CORPUS = os.path.join(HERE, "golden", "corpus")

SHAPE_SAMPLES = [
    ("flat-shapes/homestuck-beta-1", None),
    ("textured-shapes/homestuck-beta-4", ["bitmap/homestuck-beta-3"]),
    ("flat-shapes/squares", None),
    ("flat-shapes/triangle", None),
]
MORPH_SAMPLE = "flat-morph-shapes/homestuck-beta-29"
# reference ratios 0, 0.5, 1 (float) -> golden file name = ratio * 65536; MorphRatio(u16) nearest values
MORPH_RATIOS = [(0, "0.png"), (32768, "32768.png"), (65535, "65536.png")]


def load_ast(sample: str) -> dict:
    with open(os.path.join(CORPUS, sample, "ast.json")) as f:
        return json.load(f)


def load_bitmap_ast(rel: str) -> dict:
    with open(os.path.join(CORPUS, rel + ".ast.json")) as f:
        return json.load(f)


def load_golden_png(sample: str, name: str = "shape.png") -> np.ndarray:
    return np.array(PILImage.open(os.path.join(CORPUS, sample, name)).convert("RGBA"))


def read_text(sample: str, name: str) -> str:
    with open(os.path.join(CORPUS, sample, name)) as f:
        return f.read()


def fixture_canvas(tag: dict):
    b = tag["bounds"]
    x_min, x_max, y_min, y_max = b["x_min"], b["x_max"], b["y_min"], b["y_max"]
    if "morph_bounds" in tag:
        mb = tag["morph_bounds"]
        x_min, x_max = min(x_min, mb["x_min"]), max(x_max, mb["x_max"])
        y_min, y_max = min(y_min, mb["y_min"]), max(y_max, mb["y_max"])
    w = math.ceil((x_max - x_min) / 20)
    h = math.ceil((y_max - y_min) / 20)
    return w, h, [1.0, 1.0, 0.0, 0.0, float(-x_min), float(-y_min)]


# ---------------------------------------------------------------------------------------------
# a tiny scene description both sides understand
# ---------------------------------------------------------------------------------------------


class Scene:
    """width, height, definitions, bitmaps and frames of (kind, def index, matrix, ratio) items."""

    def __init__(self, width, height):
        self.width, self.height = width, height
        self.shapes = []  # define-shape ASTs
        self.morphs = []  # define-morph-shape ASTs
        self.bitmaps = {}  # id -> straight RGBA8 array
        self.frames = [[]]

    def add_shape(self, tag):
        self.shapes.append(tag)
        return len(self.shapes) - 1

    def add_morph(self, tag):
        self.morphs.append(tag)
        return len(self.morphs) - 1

    def draw_shape(self, idx, matrix, frame=0, cx=None):
        """cx (optional): colour transform of the draw, eight integers (mult x 4 in 8.8, add x 4)."""
        while len(self.frames) <= frame:
            self.frames.append([])
        self.frames[frame].append(("shape", idx, list(matrix), 0, cx))

    def draw_morph(self, idx, matrix, ratio, frame=0, ratio_f=None, cx=None):
        """ratio: MorphRatio(u16); ratio_f (optional): the TypeScript renderer's float ratio, replaces it."""
        while len(self.frames) <= frame:
            self.frames.append([])
        self.frames[frame].append(("morph", idx, list(matrix), int(ratio) if ratio_f is None else (int(ratio), float(ratio_f)), cx))


def render_oracle(scene: Scene, frame=0, want_debug=False):
    from oracle import compile_shape as cs
    from oracle import raster

    b = raster._Builder(scene.bitmaps)
    shape_defs = {}
    morph_compiled = {}
    for kind, idx, m, ratio, cx in scene.frames[frame]:
        if kind == "shape":
            if idx not in shape_defs:
                shape_defs[idx] = raster.add_shape_def(b, cs.compile_shape(scene.shapes[idx]))
            b.add_item(shape_defs[idx], m, cx=cx)
        else:
            if idx not in morph_compiled:
                morph_compiled[idx] = cs.compile_morph_shape(scene.morphs[idx])
            if isinstance(ratio, tuple):
                raster.add_morph_shape_item(b, morph_compiled[idx], m, ratio[0], ratio[1], cx=cx)
            else:
                raster.add_morph_shape_item(b, morph_compiled[idx], m, ratio, cx=cx)
    return raster.render_scene(b.scene(scene.width, scene.height), want_debug)


def make_product(scene: Scene, device=0, cuda_stream=None):
    """Registers everything of the scene in a new product renderer; returns (renderer, stages)."""
    import swf_renderer_b200 as sw

    r = sw.HeadlessRenderer(scene.width, scene.height, device=device, cuda_stream=cuda_stream)
    for bid, rgba in scene.bitmaps.items():
        r.register_bitmap(bid, rgba)
    shape_ids = [r.register_shape(t) for t in scene.shapes]
    morph_ids = [r.register_morph_shape(t) for t in scene.morphs]
    stages = []
    for items in scene.frames:
        st = sw.Stage()
        for kind, idx, m, ratio, cx in items:
            if kind == "shape":
                st.display_root.append(sw.StoredShape(shape_ids[idx], sw.Matrix2D(m), cx))
            elif isinstance(ratio, tuple):
                st.display_root.append(sw.StoredMorphShape(morph_ids[idx], sw.Matrix2D(m), ratio[0], ratio[1], cx))
            else:
                st.display_root.append(sw.StoredMorphShape(morph_ids[idx], sw.Matrix2D(m), ratio, None, cx))
        stages.append(st)
    return r, stages


def corpus_scene(sample: str, bitmaps=None) -> Scene:
    tag = load_ast(sample)
    w, h, m = fixture_canvas(tag)
    sc = Scene(w, h)
    if bitmaps:
        from oracle import decode_bitmap

        for rel in bitmaps:
            bt = load_bitmap_ast(rel)
            sc.bitmaps[bt["id"]] = decode_bitmap.define_bitmap_rgba(bt)
    sc.draw_shape(sc.add_shape(tag), m)
    return sc


def morph_scene(ratios) -> Scene:
    tag = load_ast(MORPH_SAMPLE)
    w, h, m = fixture_canvas(tag)
    sc = Scene(w, h)
    idx = sc.add_morph(tag)
    for f, r in enumerate(ratios):
        sc.draw_morph(idx, m, r, frame=f)
    return sc
