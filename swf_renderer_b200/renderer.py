"""Host-side mirror of the reference renderer interface, on top of the C ABI (include/swfr.h).

Names and argument meaning follow the reference so that tests read like its own:

  rs/src/stage.rs:4-59            Stage, DisplayPrimitive::{Shape, MorphShape}, Matrix2D, MorphRatio
  rs/src/asset.rs:9-12            ClientAssetStore.register_shape / register_morph_shape -> ShapeId / MorphShapeId
  rs/src/swf_renderer.rs:3-5      SwfRenderer.render(stage)
  rs/src/renderer.rs:89-103       Image{meta{width,height,stride}, data}
  rs/src/headless_renderer.rs     HeadlessGfxRenderer.new(w, h) / define_shape / get_image
  ts/src/lib/renderer.ts:4-8      Renderer.render(stage) / addBitmap(tag)

Errors: the reference panics / throws / returns ``Err(&'static str)``; here every failing call raises
``SwfrError`` carrying the status code of swfr.h and the library's message.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Union

import numpy as np

from . import capi
from .swf_tree import convert_define_shape


class SwfrError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("%s (status %d)" % (message, status))
        self.status = status


@dataclass
class Matrix2D:
    """rs/src/stage.rs:12-26: [scale_x, scale_y, rotate_skew0, rotate_skew1, translate_x, translate_y]."""

    c: Sequence[float] = (1.0, 1.0, 0.0, 0.0, 0.0, 0.0)

    @staticmethod
    def translate(tx: float, ty: float) -> "Matrix2D":
        return Matrix2D((1.0, 1.0, 0.0, 0.0, float(tx), float(ty)))


@dataclass
class StoredShape:  # DisplayPrimitive::Shape
    id: int
    matrix: Matrix2D = field(default_factory=Matrix2D)
    # swf-tree ColorTransformWithAlpha (not an input of the reference renderer): eight integers, red/green/blue/alpha
    # mult in Sfixed8P8 epsilons (256 = 1.0) then red/green/blue/alpha add; None = identity
    color_transform: Optional[Sequence[int]] = None


@dataclass
class StoredMorphShape:  # DisplayPrimitive::MorphShape
    id: int
    matrix: Matrix2D = field(default_factory=Matrix2D)
    ratio: int = 0  # MorphRatio(u16): 0 = start, 65535 = end
    ratio_f: Optional[float] = None  # the TypeScript renderer's ratio (0..1, float32); replaces `ratio` when set
    color_transform: Optional[Sequence[int]] = None  # see StoredShape


DisplayPrimitive = Union[StoredShape, StoredMorphShape]


@dataclass
class Stage:
    display_root: List[DisplayPrimitive] = field(default_factory=list)
    background_color: Sequence[int] = (0, 0, 0, 0)


@dataclass
class ImageMetadata:
    width: int
    height: int
    stride: int


@dataclass
class Image:
    meta: ImageMetadata
    data: np.ndarray  # (height, width, 4) uint8 RGBA


def _stage_arrays(stages: Sequence[Stage]):
    """Stage objects -> (swfr_stage[], keep-alive list)."""
    arr = (capi.Stage * len(stages))()
    keep = []
    for i, st in enumerate(stages):
        prims = (capi.DisplayPrimitive * max(1, len(st.display_root)))()
        for j, p in enumerate(st.display_root):
            prims[j].id = p.id
            prims[j].matrix[:] = [float(v) for v in p.matrix.c]
            if isinstance(p, StoredMorphShape):
                prims[j].kind = capi.PRIM_MORPH_SHAPE
                prims[j].ratio = int(p.ratio)
                if p.ratio_f is not None:
                    prims[j].flags = capi.PRIM_RATIO_F32
                    prims[j].ratio_f = float(p.ratio_f)
            else:
                prims[j].kind = capi.PRIM_SHAPE
            if p.color_transform is not None:
                prims[j].flags |= capi.PRIM_COLOR_TRANSFORM
                prims[j].color_transform = capi.ColorTransform(*[int(v) for v in p.color_transform])
        keep.append(prims)
        arr[i].background_color = capi.Rgba8(*[int(v) for v in st.background_color])
        arr[i].n_primitives = len(st.display_root)
        arr[i].display_root = C.cast(prims, C.POINTER(capi.DisplayPrimitive))
    return arr, keep


class HeadlessRenderer:
    """SwfRenderer + ClientAssetStore over libswfr_b200 (one CUDA device, one stream per instance)."""

    def __init__(self, width: int, height: int, device: int = 0, cuda_stream: Optional[int] = None):
        self._lib = capi.load()
        self._h = C.c_void_p()
        self.width, self.height = int(width), int(height)
        self.device = int(device)
        if cuda_stream is None:
            rc = self._lib.swfr_create(device, self.width, self.height, C.byref(self._h))
        else:
            rc = self._lib.swfr_create_on_stream(device, self.width, self.height, C.c_void_p(cuda_stream), C.byref(self._h))
        if rc != capi.OK:
            self._h = C.c_void_p()
            raise SwfrError(rc, "swfr_create failed: %s" % self._lib.swfr_status_string(rc).decode())

    # -- lifecycle ------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.swfr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != capi.OK:
            raise SwfrError(rc, self._lib.swfr_last_error(self._h).decode() or self._lib.swfr_status_string(rc).decode())

    def set_option(self, key: int, value: int):
        self._check(self._lib.swfr_set_option(self._h, key, value))

    # -- ClientAssetStore -----------------------------------------------------------------------------
    def register_shape(self, tag) -> int:
        """``tag``: define-shape AST dict (swf-tree JSON) or a converted capi.DefineShape."""
        cv = convert_define_shape(tag) if isinstance(tag, dict) else None
        out = C.c_uint32()
        self._check(self._lib.swfr_register_shape(self._h, C.byref(cv.tag if cv else tag), C.byref(out)))
        return out.value

    define_shape = register_shape  # HeadlessGfxRenderer::define_shape

    def register_morph_shape(self, tag) -> int:
        cv = convert_define_shape(tag) if isinstance(tag, dict) else None
        out = C.c_uint32()
        self._check(self._lib.swfr_register_morph_shape(self._h, C.byref(cv.tag if cv else tag), C.byref(out)))
        return out.value

    def add_bitmap(self, tag: dict):
        """Renderer.addBitmap(tag: DefineBitmap) (ts/src/lib/renderer.ts:7)."""
        if tag["media_type"] != "image/x-swf-bmp":
            raise SwfrError(capi.ERR_UNSUPPORTED_STYLE, "NotImplemented: Support for %s images" % tag["media_type"])
        data = bytes.fromhex(tag["data"])
        self._check(self._lib.swfr_register_bitmap_xswfbmp(self._h, tag["id"], data, len(data)))

    def register_bitmap(self, bitmap_id: int, rgba: np.ndarray):
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        h, w = rgba.shape[:2]
        self._check(self._lib.swfr_register_bitmap(self._h, bitmap_id, w, h, rgba.ctypes.data, w * 4))

    # -- SwfRenderer ----------------------------------------------------------------------------------
    def render(self, stage: Stage):
        arr, keep = _stage_arrays([stage])
        self._check(self._lib.swfr_render(self._h, arr))

    def render_batch(self, stages: Sequence[Stage]):
        arr, keep = _stage_arrays(stages)
        self._check(self._lib.swfr_render_batch(self._h, arr, len(stages)))

    def sync(self):
        self._check(self._lib.swfr_sync(self._h))

    def get_image(self, frame: int = 0, premultiplied: bool = False) -> Image:
        """HeadlessGfxRenderer::get_image / download_image.  Straight alpha by default (PNG-export rounding)."""
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        self._check(self._lib.swfr_read_image(self._h, frame, out.ctypes.data, self.width * 4, 1 if premultiplied else 0))
        return Image(ImageMetadata(self.width, self.height, self.width * 4), out)

    def device_frames(self):
        """The finished frames of the last render where they lie in HBM, as a torch uint8 tensor
        [frames, height, width, 4] (premultiplied RGBA8, no copy; valid until the next render).  For callers that keep
        the pixels on the GPU or move them between GPUs (`sharding.gather_frames`); call sync() first."""
        import torch

        ptr, n = C.c_void_p(), C.c_uint32()
        self._check(self._lib.swfr_device_frames(self._h, C.byref(ptr), C.byref(n)))
        shape = (int(n.value), self.height, self.width, 4)

        class _Frames:  # minimal CUDA array interface holder
            __cuda_array_interface__ = {"shape": shape, "typestr": "|u1", "data": (int(ptr.value or 0), False), "version": 2}

        if n.value == 0 or not ptr.value:
            return torch.empty((0, self.height, self.width, 4), dtype=torch.uint8, device="cuda:%d" % self.device)
        return torch.as_tensor(_Frames(), device="cuda:%d" % self.device)

    # -- optional gather of finished frames onto one GPU (off the hot path) --------------------
    def export_frames(self) -> bytes:
        """swfr_export_frames: the bytes another renderer (same or other process) needs to copy this renderer's
        finished frames GPU to GPU.  Waits for the last render."""
        exp = capi.FramesExport()
        self._check(self._lib.swfr_export_frames(self._h, C.byref(exp)))
        return bytes(exp)

    def gather_frames(self, exports: Sequence[bytes]):
        """swfr_gather_frames: copies the frames of every exporting renderer (in rank order; frame f was rendered by
        renderer f mod N in slot f div N) into this renderer's gather buffer, in global frame order - one strided
        asynchronous peer copy per source on the copy stream.  Returns (torch uint8 tensor [frames, H, W, 4] over the
        gather buffer, device milliseconds of the copies)."""
        import torch

        arr = (capi.FramesExport * len(exports))()
        for k, b in enumerate(exports):
            C.memmove(C.byref(arr[k]), b, C.sizeof(capi.FramesExport))
        ptr, n = C.c_void_p(), C.c_uint32()
        self._check(self._lib.swfr_gather_frames(self._h, arr, len(exports), C.byref(ptr), C.byref(n)))
        ms = C.c_float()
        self._check(self._lib.swfr_gather_last_ms(self._h, C.byref(ms)))  # waits for the copies
        shape = (int(n.value), self.height, self.width, 4)

        class _Frames:
            __cuda_array_interface__ = {"shape": shape, "typestr": "|u1", "data": (int(ptr.value or 0), False), "version": 2}

        return torch.as_tensor(_Frames(), device="cuda:%d" % self.device), float(ms.value)

    def stats(self) -> dict:
        st = capi.Stats()
        self._check(self._lib.swfr_get_stats(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in capi.Stats._fields_}

    def stage_times(self) -> dict:
        """Per-stage device milliseconds of the last render (needs set_option(OPT_PROFILE, 1) before it)."""
        ms = (C.c_float * 16)()
        n, passes = C.c_uint32(), C.c_uint32()
        self._check(self._lib.swfr_get_stage_times(self._h, ms, 16, C.byref(n), C.byref(passes)))
        names = [self._lib.swfr_stage_name(i).decode() for i in range(n.value)]
        return {"passes": passes.value, "ms": {names[i]: float(ms[i]) for i in range(n.value)}}

    # -- stages resident in HBM (repeated rendering without host traffic) -------------------------------
    def create_batch(self, stages) -> "ResidentBatch":
        """``stages``: sequence of Stage objects, or a prebuilt (capi.Stage array, keep-alive) pair."""
        arr, keep = stages if isinstance(stages, tuple) else _stage_arrays(stages)
        h = C.c_void_p()
        self._check(self._lib.swfr_batch_create(self._h, arr, len(arr), C.byref(h)))
        return ResidentBatch(self, h, len(arr))

    def render_stage_array(self, arr, n: int):
        """swfr_render_batch on a prebuilt capi.Stage array (host buffers; used by the end-to-end benchmark)."""
        self._check(self._lib.swfr_render_batch(self._h, arr, n))

    def read_frames_async(self, first: int, count: int, dst_ptr: int):
        self._check(self._lib.swfr_read_frames_async(self._h, first, count, C.c_void_p(dst_ptr)))

    # -- parity taps ----------------------------------------------------------------------------------
    def debug_compiled(self, kind: int, def_id: int):
        nc, npth = C.c_uint64(), C.c_uint64()
        self._check(self._lib.swfr_debug_compiled(self._h, kind, def_id, None, 0, C.byref(nc), None, 0, C.byref(npth)))
        cmds = np.zeros((max(1, nc.value), 9), dtype=np.float64)
        info = np.zeros((max(1, npth.value), 3), dtype=np.int32)
        self._check(
            self._lib.swfr_debug_compiled(
                self._h, kind, def_id, cmds.ctypes.data, nc.value, C.byref(nc), info.ctypes.data, npth.value, C.byref(npth)
            )
        )
        return cmds[: nc.value], info[: npth.value]

    def debug_segments(self, kind: int, def_id: int) -> np.ndarray:
        n = C.c_uint64()
        self._check(self._lib.swfr_debug_segments(self._h, kind, def_id, None, 0, C.byref(n)))
        segs = np.zeros((max(1, n.value), 14), dtype=np.float64)
        self._check(self._lib.swfr_debug_segments(self._h, kind, def_id, segs.ctypes.data, n.value, C.byref(n)))
        return segs[: n.value]

    def debug_edges(self, frame: int = 0):
        n = C.c_uint64()
        self._check(self._lib.swfr_debug_edges(self._h, frame, None, None, 0, C.byref(n)))
        edges = np.zeros((max(1, n.value), 4), dtype=np.int32)
        epath = np.zeros(max(1, n.value), dtype=np.int32)
        self._check(self._lib.swfr_debug_edges(self._h, frame, edges.ctypes.data, epath.ctypes.data, n.value, C.byref(n)))
        return edges[: n.value], epath[: n.value]

    def debug_tile_counts(self, frame: int = 0) -> np.ndarray:
        ty, tx = (self.height + 15) // 16, (self.width + 15) // 16
        out = np.zeros((ty, tx), dtype=np.uint32)
        self._check(self._lib.swfr_debug_tile_counts(self._h, frame, out.ctypes.data, out.size))
        return out


class ResidentBatch:
    """A set of stages flattened and uploaded once (swfr_batch_create)."""

    def __init__(self, renderer: HeadlessRenderer, handle, n_frames: int):
        self._r, self._h, self.n_frames = renderer, handle, n_frames

    def render(self):
        self._r._check(self._r._lib.swfr_batch_render(self._r._h, self._h))

    def close(self):
        if self._h and self._h.value and self._r._h.value:
            self._r._lib.swfr_batch_destroy(self._r._h, self._h)
        self._h = C.c_void_p()


def stage_array_from_numpy(ids: np.ndarray, matrices: np.ndarray, kinds=None, ratios=None):
    """Bulk construction of ONE swfr_stage from arrays (ids[n], matrices[n,6] float32) without Python loops."""
    n = len(ids)
    prims = (capi.DisplayPrimitive * max(1, n))()
    P = capi.DisplayPrimitive
    dt = np.dtype(
        {
            "names": ["kind", "id", "matrix", "ratio"],
            "formats": [np.uint32, np.uint32, (np.float32, 6), np.uint16],
            "offsets": [P.kind.offset, P.id.offset, P.matrix.offset, P.ratio.offset],
            "itemsize": C.sizeof(P),
        }
    )
    v = np.frombuffer(prims, dtype=dt)
    v["id"][:n] = ids
    v["matrix"][:n] = matrices
    v["kind"][:n] = 0 if kinds is None else kinds
    v["ratio"][:n] = 0 if ratios is None else ratios
    return prims, n


def stages_from_prims(prim_arrays):
    """[(DisplayPrimitive array, n)] -> (capi.Stage array, keep-alive)."""
    arr = (capi.Stage * len(prim_arrays))()
    for i, (prims, n) in enumerate(prim_arrays):
        arr[i].n_primitives = n
        arr[i].display_root = C.cast(prims, C.POINTER(capi.DisplayPrimitive))
    return arr, list(prim_arrays)


def decode_x_swf_bmp(data: bytes) -> np.ndarray:
    """decodeXSwfBmpSync through the library's host decoder (no GPU needed)."""
    lib = capi.load()
    w, h = C.c_uint32(), C.c_uint32()
    rc = lib.swfr_decode_xswfbmp(data, len(data), None, 0, C.byref(w), C.byref(h))
    if rc != capi.OK:
        raise SwfrError(rc, lib.swfr_status_string(rc).decode())
    out = np.empty((h.value, w.value, 4), dtype=np.uint8)
    rc = lib.swfr_decode_xswfbmp(data, len(data), out.ctypes.data, out.size, C.byref(w), C.byref(h))
    if rc != capi.OK:
        raise SwfrError(rc, lib.swfr_status_string(rc).decode())
    return out


def compile_tag(tag: dict, morph: bool = False):
    """Host-only shape compile (decodeSwfShape / decodeSwfMorphShape restated in the library).

    Returns (commands[n,9], path_info[n_paths,3], segments[n_seg,14]); see swfr_debug_compiled in swfr.h."""
    import numpy as _np

    lib = capi.load()
    cv = convert_define_shape(tag)
    nc, npth, ns = C.c_uint64(), C.c_uint64(), C.c_uint64()
    rc = lib.swfr_compile_debug(C.byref(cv.tag), int(morph), None, 0, C.byref(nc), None, 0, C.byref(npth), None, 0, C.byref(ns))
    if rc != capi.OK:
        raise SwfrError(rc, lib.swfr_status_string(rc).decode())
    cmds = _np.zeros((max(1, nc.value), 9), dtype=_np.float64)
    info = _np.zeros((max(1, npth.value), 3), dtype=_np.int32)
    segs = _np.zeros((max(1, ns.value), 14), dtype=_np.float64)
    rc = lib.swfr_compile_debug(
        C.byref(cv.tag), int(morph), cmds.ctypes.data, nc.value, C.byref(nc), info.ctypes.data, npth.value, C.byref(npth),
        segs.ctypes.data, ns.value, C.byref(ns),
    )
    if rc != capi.OK:
        raise SwfrError(rc, lib.swfr_status_string(rc).decode())
    return cmds[: nc.value], info[: npth.value], segs[: ns.value]


def debug_morph_stroke(tag: dict, ratio: float):
    """Host run of the device stroker's generator (swfr_debug_morph_stroke) for one DefineMorphShape tag at `ratio`.

    Returns (segments[n, 8] = curve flag, line path index, x0, y0, cx, cy, x1, y1 in twips; number of visible line paths)."""
    import numpy as _np

    lib = capi.load()
    cv = convert_define_shape(tag)
    n, paths = C.c_uint64(), C.c_uint32()
    rc = lib.swfr_debug_morph_stroke(C.byref(cv.tag), float(ratio), None, 0, C.byref(n), C.byref(paths))
    if rc != capi.OK:
        raise SwfrError(rc, lib.swfr_status_string(rc).decode())
    segs = _np.zeros((max(1, n.value), 8), dtype=_np.float64)
    rc = lib.swfr_debug_morph_stroke(C.byref(cv.tag), float(ratio), segs.ctypes.data, n.value, C.byref(n), C.byref(paths))
    if rc != capi.OK:
        raise SwfrError(rc, lib.swfr_status_string(rc).decode())
    return segs[: n.value], int(paths.value)
