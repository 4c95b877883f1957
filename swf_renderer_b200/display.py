"""Mirror of the TypeScript renderer's public surface over the C ABI:

  ts/src/lib/display/stage.ts:7-18                 Stage {backgroundColor?, width, height, children}
  ts/src/lib/display/display-object-container.ts   DisplayObjectContainer {children, matrix?}
  ts/src/lib/display/shape.ts / morph-shape.ts     Shape {definition, matrix?} / MorphShape {definition, matrix?, ratio}
  ts/src/lib/renderer.ts:4-8                       Renderer {render(stage), addBitmap(tag)}
  ts/src/lib/renderers/canvas-renderer.ts:61-145   render / drawDisplayObject / drawContainer / definition caches
  ts/src/lib/image-data-to-pam.ts:8-30             imageDataToPam

Definitions are swf-tree JSON dicts (as in the reference's tests); like the reference, a definition is compiled the
first time it is drawn and cached per object (canvas-renderer.ts:96-112 uses WeakMaps keyed by the tag object).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Union

import numpy as np

from . import capi
from .renderer import HeadlessRenderer, Image, SwfrError


@dataclass
class Matrix:
    """swf-tree Matrix: Sfixed16P16 epsilons (scale / rotate-skew) and twips (translate)."""
    scale_x: int = 65536
    scale_y: int = 65536
    rotate_skew0: int = 0
    rotate_skew1: int = 0
    translate_x: int = 0
    translate_y: int = 0


@dataclass
class ColorTransform:
    """swf-tree ColorTransformWithAlpha: mult in Sfixed8P8 epsilons (256 = 1.0), add in integers.  The reference's
    display objects carry no colour transform; this is the SWF place-object field, concatenated down the tree."""
    red_mult: int = 256
    green_mult: int = 256
    blue_mult: int = 256
    alpha_mult: int = 256
    red_add: int = 0
    green_add: int = 0
    blue_add: int = 0
    alpha_add: int = 0


@dataclass
class Shape:
    definition: dict
    matrix: Optional[Matrix] = None
    color_transform: Optional[ColorTransform] = None


@dataclass
class MorphShape:
    definition: dict
    matrix: Optional[Matrix] = None
    ratio: float = 0.0  # 0..1
    color_transform: Optional[ColorTransform] = None


@dataclass
class DisplayObjectContainer:
    children: List["DisplayObject"] = field(default_factory=list)
    matrix: Optional[Matrix] = None
    color_transform: Optional[ColorTransform] = None


DisplayObject = Union[DisplayObjectContainer, MorphShape, Shape]


@dataclass
class Stage:
    width: int
    height: int
    children: List[DisplayObject] = field(default_factory=list)
    background_color: Optional[Sequence[int]] = None  # StraightSRgba8 (r, g, b, a)


def _matrix(m: Matrix) -> capi.SwfMatrix:
    return capi.SwfMatrix(int(m.scale_x), int(m.scale_y), int(m.rotate_skew0), int(m.rotate_skew1), int(m.translate_x),
                          int(m.translate_y))


class CanvasRenderer:
    """`Renderer` of ts/src/lib/renderer.ts on a CUDA device: render(stage), add_bitmap(tag), plus the image export the
    reference's test harness performs on its canvas (toBuffer("image/png"), imageDataToPam)."""

    def __init__(self, width: int, height: int, device: int = 0):
        self.width, self.height = width, height
        self._r = HeadlessRenderer(width, height, device=device)
        self._shape_cache = {}  # id(definition) -> (definition, ShapeId)   (keeps the tag alive, like a WeakMap entry)
        self._morph_cache = {}

    # -- Renderer ---------------------------------------------------------------------------------------
    def add_bitmap(self, tag: dict) -> None:
        """addBitmap(tag: DefineBitmap) - node-canvas-bitmap-service.ts:14-37 (only image/x-swf-bmp is implemented
        there; other media types throw NotImplementedBitmapType)."""
        self._r.add_bitmap(tag)

    def render(self, stage: Stage) -> None:
        self.render_batch([stage])

    def render_batch(self, stages: Sequence[Stage]) -> None:
        arr = (capi.DisplayStage * len(stages))()
        keep = []
        for i, st in enumerate(stages):
            if st.width != self.width or st.height != self.height:
                raise ValueError("stage size differs from the renderer's viewport")
            arr[i].width, arr[i].height = st.width, st.height
            if st.background_color is not None:
                arr[i].has_background_color = 1
                arr[i].background_color = capi.Rgba8(*[int(v) for v in st.background_color])
            kids = self._objects(st.children, keep)
            arr[i].n_children = len(st.children)
            arr[i].children = C.cast(kids, C.POINTER(capi.DisplayObject))
        self._r._check(self._r._lib.swfr_render_display_stages(self._r._h, arr, len(stages)))

    # -- export -----------------------------------------------------------------------------------------
    def get_image(self, frame: int = 0, premultiplied: bool = False) -> Image:
        return self._r.get_image(frame=frame, premultiplied=premultiplied)

    def to_png(self, frame: int = 0) -> bytes:
        return write_png(self.get_image(frame).data)

    def to_pam(self, frame: int = 0) -> bytes:
        return image_data_to_pam(self.get_image(frame).data)

    def set_option(self, key: int, value: int) -> None:
        self._r.set_option(key, value)

    def stats(self):
        return self._r.stats()

    def close(self):
        self._r.close()

    # -- internals --------------------------------------------------------------------------------------
    def _objects(self, objs: Sequence[DisplayObject], keep: list):
        arr = (capi.DisplayObject * max(1, len(objs)))()
        keep.append(arr)
        for i, o in enumerate(objs):
            if o.matrix is not None:
                arr[i].has_matrix = 1
                arr[i].matrix = _matrix(o.matrix)
            if o.color_transform is not None:
                t = o.color_transform
                arr[i].has_color_transform = 1
                arr[i].color_transform = capi.ColorTransform(t.red_mult, t.green_mult, t.blue_mult, t.alpha_mult, t.red_add,
                                                             t.green_add, t.blue_add, t.alpha_add)
            if isinstance(o, DisplayObjectContainer):
                arr[i].type = capi.DISPLAY_CONTAINER
                kids = self._objects(o.children, keep)
                arr[i].n_children = len(o.children)
                arr[i].children = C.cast(kids, C.POINTER(capi.DisplayObject))
            elif isinstance(o, MorphShape):
                arr[i].type = capi.DISPLAY_MORPH_SHAPE
                arr[i].id = self._compiled(self._morph_cache, o.definition, True)
                arr[i].ratio = float(o.ratio)
            elif isinstance(o, Shape):
                arr[i].type = capi.DISPLAY_SHAPE
                arr[i].id = self._compiled(self._shape_cache, o.definition, False)
            else:
                raise TypeError("UnexpectedDisplayObjectType")  # canvas-renderer.ts:91-92
        return arr

    def _compiled(self, cache: dict, definition: dict, morph: bool) -> int:
        hit = cache.get(id(definition))
        if hit is None:
            rid = self._r.register_morph_shape(definition) if morph else self._r.register_shape(definition)
            cache[id(definition)] = hit = (definition, rid)
        return hit[1]


def flatten_stage(stage_struct) -> np.ndarray:
    """swfr_flatten_display_stage on a capi.DisplayStage (host only): structured array of the primitives."""
    lib = capi.load()
    n = C.c_uint32()
    rc = lib.swfr_flatten_display_stage(C.byref(stage_struct), None, 0, C.byref(n))
    if rc != capi.OK:
        raise SwfrError(rc, lib.swfr_status_string(rc).decode())
    prims = (capi.DisplayPrimitive * max(1, n.value))()
    rc = lib.swfr_flatten_display_stage(C.byref(stage_struct), prims, n.value, C.byref(n))
    if rc != capi.OK:
        raise SwfrError(rc, lib.swfr_status_string(rc).decode())
    return prims, n.value


def _write(fn_name: str, rgba: np.ndarray) -> bytes:
    lib = capi.load()
    img = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = img.shape[:2]
    n = C.c_uint64()
    fn = getattr(lib, fn_name)
    rc = fn(img.ctypes.data, w, h, w * 4, None, 0, C.byref(n))
    if rc != capi.OK:
        raise SwfrError(rc, lib.swfr_status_string(rc).decode())
    out = (C.c_uint8 * n.value)()
    rc = fn(img.ctypes.data, w, h, w * 4, out, n.value, C.byref(n))
    if rc != capi.OK:
        raise SwfrError(rc, lib.swfr_status_string(rc).decode())
    return bytes(out)


def image_data_to_pam(rgba: np.ndarray) -> bytes:
    """imageDataToPam (ts/src/lib/image-data-to-pam.ts:8-30) / write_pam (rs/src/pam.rs:3-34)."""
    return _write("swfr_write_pam", rgba)


def write_png(rgba: np.ndarray) -> bytes:
    """canvas.toBuffer("image/png") of the reference test harness: 8-bit straight-alpha RGBA."""
    return _write("swfr_write_png", rgba)
