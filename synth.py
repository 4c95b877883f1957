"""Synthetic SWF shape streams, shared by bench.py and the parity tests.

One frame = ``n_shapes`` random star-shaped DefineShape tags + one display primitive each, from a counter-based
splitmix64 stream seeded with ``0x5EED0000 + frame_index`` (value k of shape i = mix(seed + (256 i + k + 1) * GOLDEN),
so the stream is reproducible from any language).  Per shape:
  centre uniform over the frame; circum-radius log-uniform in [4, 256] px (x ``radius_scale``);
  n in U{3..32} vertices at jittered angles, radius jitter +-40 %, integer twips;
  each edge curved with probability 0.5 (control = chord midpoint pushed +-25 % of the chord along its normal);
  20 % carry an inner contour with a second fill (inner edges have BOTH a left and a right fill);
  fills: 60 % solid, 10 % linear, 10 % radial, 10 % focal, 10 % bitmap (one of 8 seeded 256x256 textures, half repeating);
  alpha 255 with probability 0.5 else U[32, 254]; placement = translate (80 %) or rotate + scale in [0.5, 2] (20 %).
Everything is generated with numpy for all shapes at once; ``ast(i)`` gives the swf-tree JSON form of shape i
(what the oracle and the dict-based binding consume) and ``register(renderer)`` feeds the C ABI directly.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
SLOTS = 256
MAXV = 32
N_TEXTURES = 8


def _mix(z):
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _u(seed: int, shape: np.ndarray, slot) -> np.ndarray:
    """uniform [0,1) doubles for (shape index, slot)."""
    with np.errstate(over="ignore"):
        idx = shape.astype(np.uint64) * np.uint64(SLOTS) + np.asarray(slot, dtype=np.uint64) + np.uint64(1)
        z = np.uint64(seed) + idx * GOLDEN
    return (_mix(z) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def textures(seed: int = 0x7E87) -> list:
    """8 seeded 256x256 straight-RGBA textures: value noise over a checker; odd ones carry alpha."""
    out = []
    yy, xx = np.mgrid[0:256, 0:256]
    for t in range(N_TEXTURES):
        n = _u(seed + t, (yy * 256 + xx).ravel(), 0).reshape(256, 256)
        n2 = _u(seed + t, (yy * 256 + xx).ravel(), 1).reshape(256, 256)
        check = (((xx >> (3 + t % 3)) + (yy >> (3 + t % 3))) & 1).astype(np.float64)
        img = np.zeros((256, 256, 4), dtype=np.uint8)
        img[..., 0] = np.clip(255 * (0.6 * check + 0.4 * n), 0, 255)
        img[..., 1] = np.clip(255 * (0.5 * n2 + 0.5 * (xx / 255.0)), 0, 255)
        img[..., 2] = np.clip(255 * (0.7 * (1 - check) + 0.3 * (yy / 255.0)), 0, 255)
        img[..., 3] = 255 if t % 2 == 0 else np.clip(64 + 191 * n, 0, 255).astype(np.uint8)
        out.append(img)
    return out


class SynthFrame:
    def __init__(self, frame_index: int, n_shapes: int, width: int, height: int, radius_scale: float = 1.0,
                 solid_only: bool = False, stroke_fraction: float = 0.0):
        self.seed = 0x5EED0000 + frame_index
        self.n, self.width, self.height = n_shapes, width, height
        s = np.arange(n_shapes)
        u = lambda slot: _u(self.seed, s, slot)
        self.cx = u(0) * width
        self.cy = u(1) * height
        self.radius = 4.0 * np.power(64.0, u(2)) * radius_scale  # px
        self.nv = 3 + np.floor(u(3) * 30).astype(np.int64)
        self.inner = u(4) < 0.2
        ft = u(5)
        # 0 solid, 1 linear, 2 radial, 3 focal, 4 bitmap
        self.fill_type = np.select([ft < 0.6, ft < 0.7, ft < 0.8, ft < 0.9], [0, 1, 2, 3], 4)
        if solid_only:
            self.fill_type[:] = 0
        self.color = np.stack([np.floor(u(6 + c) * 256) for c in range(3)], axis=1).astype(np.int64)
        self.alpha = np.where(u(9) < 0.5, 255, 32 + np.floor(u(10) * 223)).astype(np.int64)
        self.color2 = np.stack([np.floor(u(11 + c) * 256) for c in range(3)], axis=1).astype(np.int64)
        self.alpha2 = np.where(u(14) < 0.5, 255, 32 + np.floor(u(15) * 223)).astype(np.int64)
        self.n_stops = 2 + np.floor(u(16) * 3).astype(np.int64)  # 2..4
        self.spread = np.floor(u(17) * 3).astype(np.int64)
        self.linear_rgb = u(18) < 0.1
        self.focal = np.floor((u(19) * 1.8 - 0.9) * 256).astype(np.int64)  # Sfixed8P8 epsilons
        self.grad_rot = u(20) * 2 * math.pi
        self.tex = np.floor(u(21) * N_TEXTURES).astype(np.int64)
        self.tex_repeat = u(22) < 0.5
        self.tex_scale = np.select([u(23) < 0.33, u(23) < 0.66], [0.5, 1.0], 2.58)
        self.rotated = u(24) < 0.2
        self.rot = u(25) * 2 * math.pi
        self.scale = 0.5 + 1.5 * u(26)
        # SURVEY 8d config 5, "separate variant": 2 px strokes (opaque, the second colour) on this fraction of the shapes
        self.stroked = u(27) < stroke_fraction
        self.stop_colors = np.stack(
            [np.stack([np.floor(u(32 + 4 * k + c) * 256) for c in range(4)], axis=1) for k in range(4)], axis=1
        ).astype(np.int64)  # (n, 4 stops, rgba)
        opaque_ramp = u(48) < 0.5
        self.stop_colors[opaque_ramp, :, 3] = 255
        # ---- geometry (twips, shape-local, centred on the origin) ----
        k = np.arange(MAXV)[None, :]
        ss = s[:, None]
        ang = 2 * math.pi * (k + 0.5 + 0.8 * (_u(self.seed, ss, 64 + 4 * k) - 0.5)) / self.nv[:, None]
        rad = self.radius[:, None] * 20.0 * (1.0 + 0.8 * (_u(self.seed, ss, 65 + 4 * k) - 0.5))
        self.vx = np.rint(rad * np.cos(ang)).astype(np.int64)
        self.vy = np.rint(rad * np.sin(ang)).astype(np.int64)
        self.curved = _u(self.seed, ss, 66 + 4 * k) < 0.5
        self.push = (_u(self.seed, ss, 67 + 4 * k) - 0.5) * 0.5  # +-25 % of the chord
        self.mask = k < self.nv[:, None]
        nxt = (k + 1) % self.nv[:, None]
        rows = np.arange(n_shapes)[:, None]
        self.dx = self.vx[rows, nxt] - self.vx
        self.dy = self.vy[rows, nxt] - self.vy
        # control point relative to the edge start: half the chord + push along the normal
        self.cdx = np.rint(self.dx * 0.5 - self.dy * self.push).astype(np.int64)
        self.cdy = np.rint(self.dy * 0.5 + self.dx * self.push).astype(np.int64)
        # inner contour: the outer one scaled by 1/2 (integer twips), same direction
        self.ivx = np.rint(self.vx * 0.5).astype(np.int64)
        self.ivy = np.rint(self.vy * 0.5).astype(np.int64)
        self.idx_ = self.ivx[rows, nxt] - self.ivx
        self.idy_ = self.ivy[rows, nxt] - self.ivy
        self.icdx = np.rint(self.idx_ * 0.5 - self.idy_ * self.push).astype(np.int64)
        self.icdy = np.rint(self.idy_ * 0.5 + self.idx_ * self.push).astype(np.int64)

    # ---- placement ------------------------------------------------------------------------------------
    def matrix(self, i: int):
        """Matrix2D order [scale_x, scale_y, rotate_skew0, rotate_skew1, tx, ty] as float32-exact values."""
        tx, ty = self.cx[i] * 20.0, self.cy[i] * 20.0
        if self.rotated[i]:
            c, s_ = math.cos(self.rot[i]) * self.scale[i], math.sin(self.rot[i]) * self.scale[i]
            m = [c, c, s_, -s_, tx, ty]
        else:
            m = [1.0, 1.0, 0.0, 0.0, tx, ty]
        return [float(np.float32(v)) for v in m]

    def matrices(self) -> np.ndarray:
        m = np.zeros((self.n, 6), dtype=np.float32)
        c = np.where(self.rotated, np.cos(self.rot) * self.scale, 1.0)
        s_ = np.where(self.rotated, np.sin(self.rot) * self.scale, 0.0)
        m[:, 0] = c
        m[:, 1] = c
        m[:, 2] = s_
        m[:, 3] = -s_
        m[:, 4] = self.cx * 20.0
        m[:, 5] = self.cy * 20.0
        return m

    # ---- styles ---------------------------------------------------------------------------------------
    def _fill_ast(self, i: int, second: bool):
        if second:
            c, a = self.color2[i], int(self.alpha2[i])
            return {"type": "solid", "color": {"r": int(c[0]), "g": int(c[1]), "b": int(c[2]), "a": a}}
        t = int(self.fill_type[i])
        c, a = self.color[i], int(self.alpha[i])
        if t == 0:
            return {"type": "solid", "color": {"r": int(c[0]), "g": int(c[1]), "b": int(c[2]), "a": a}}
        r_tw = float(self.radius[i]) * 20.0
        if t == 4:
            sc = int(round(20.0 * float(self.tex_scale[i]) * 65536))
            return {
                "type": "bitmap",
                "bitmap_id": int(self.tex[i]),
                "matrix": {"scale_x": sc, "scale_y": sc, "rotate_skew0": 0, "rotate_skew1": 0,
                           "translate_x": int(-r_tw), "translate_y": int(-r_tw)},
                "repeating": bool(self.tex_repeat[i]),
                "smoothed": True,
            }
        g = r_tw / 16384.0
        cs_, sn = math.cos(self.grad_rot[i]) * g, math.sin(self.grad_rot[i]) * g
        ns = int(self.n_stops[i])
        stops = []
        for k in range(ns):
            sc_ = self.stop_colors[i, k]
            stops.append({"ratio": int(round(255 * k / (ns - 1))),
                          "color": {"r": int(sc_[0]), "g": int(sc_[1]), "b": int(sc_[2]), "a": int(sc_[3])}})
        out = {
            "type": ["", "linear-gradient", "radial-gradient", "focal-gradient"][t],
            "matrix": {"scale_x": int(round(cs_ * 65536)), "scale_y": int(round(cs_ * 65536)),
                       "rotate_skew0": int(round(sn * 65536)), "rotate_skew1": int(round(-sn * 65536)),
                       "translate_x": 0, "translate_y": 0},
            "gradient": {"spread": ["pad", "reflect", "repeat"][int(self.spread[i])],
                         "color_space": "linear-rgb" if self.linear_rgb[i] else "s-rgb", "colors": stops},
        }
        if t == 3:
            out["focal_point"] = int(self.focal[i])
        return out

    def _line_ast(self, i: int):
        c = self.color2[i]
        return {"width": 40, "fill": {"type": "solid", "color": {"r": int(c[0]), "g": int(c[1]), "b": int(c[2]), "a": 255}}}

    def ast(self, i: int) -> dict:
        """swf-tree JSON form of shape i (define-shape)."""
        n = int(self.nv[i])
        fills = [self._fill_ast(i, False)]
        recs = [{"type": "style-change", "move_to": {"x": int(self.vx[i, 0]), "y": int(self.vy[i, 0])}, "left_fill": 1}]
        lines = []
        if self.stroked[i]:
            lines.append(self._line_ast(i))
            recs[0]["line_style"] = 1
        for k in range(n):
            e = {"type": "edge", "delta": {"x": int(self.dx[i, k]), "y": int(self.dy[i, k])}}
            if self.curved[i, k]:
                e["control_delta"] = {"x": int(self.cdx[i, k]), "y": int(self.cdy[i, k])}
            recs.append(e)
        if self.inner[i]:
            fills.append(self._fill_ast(i, True))
            recs.append({"type": "style-change", "move_to": {"x": int(self.ivx[i, 0]), "y": int(self.ivy[i, 0])},
                         "left_fill": 2, "right_fill": 1})
            for k in range(n):
                e = {"type": "edge", "delta": {"x": int(self.idx_[i, k]), "y": int(self.idy_[i, k])}}
                if self.curved[i, k]:
                    e["control_delta"] = {"x": int(self.icdx[i, k]), "y": int(self.icdy[i, k])}
                recs.append(e)
        r = int(self.radius[i] * 20 * 1.5) + 1
        return {
            "type": "define-shape",
            "id": i & 0xFFFF,
            "bounds": {"x_min": -r, "x_max": r, "y_min": -r, "y_max": r},
            "shape": {"initial_styles": {"fill": fills, "line": lines}, "records": recs},
        }

    # ---- fast path into the C ABI -----------------------------------------------------------------------
    def register(self, renderer) -> np.ndarray:
        """Registers every shape of the frame through swfr_register_shape; returns the ShapeIds."""
        from swf_renderer_b200 import capi
        from swf_renderer_b200.swf_tree import _fill

        n = self.n
        per = 1 + self.nv + np.where(self.inner, 1 + self.nv, 0)
        base = np.concatenate([[0], np.cumsum(per)])
        total = int(base[-1])
        recs = (capi.ShapeRecord * total)()
        R = capi.ShapeRecord
        fields = ["type", "delta_x", "delta_y", "control_delta_x", "control_delta_y", "has_control_delta", "has_move_to",
                  "has_left_fill", "has_right_fill", "has_line_style", "move_to_x", "move_to_y", "left_fill", "right_fill",
                  "line_style"]
        np_t = {1: np.uint8, 4: np.int32}
        dt = np.dtype({
            "names": fields,
            "formats": [np_t[getattr(R, f).size] for f in fields],
            "offsets": [getattr(R, f).offset for f in fields],
            "itemsize": C.sizeof(R),
        })
        v = np.frombuffer(recs, dtype=dt)
        b0 = base[:-1]
        # outer contour
        v["type"][b0] = capi.RECORD_STYLE_CHANGE
        v["has_move_to"][b0] = 1
        v["move_to_x"][b0] = self.vx[:, 0]
        v["move_to_y"][b0] = self.vy[:, 0]
        v["has_left_fill"][b0] = 1
        v["left_fill"][b0] = 1
        v["has_line_style"][b0] = self.stroked
        v["line_style"][b0] = self.stroked.astype(np.int32)
        k = np.arange(MAXV)[None, :]
        dest = (b0[:, None] + 1 + k)[self.mask]
        v["type"][dest] = capi.RECORD_EDGE
        v["delta_x"][dest] = self.dx[self.mask]
        v["delta_y"][dest] = self.dy[self.mask]
        v["has_control_delta"][dest] = self.curved[self.mask]
        v["control_delta_x"][dest] = self.cdx[self.mask]
        v["control_delta_y"][dest] = self.cdy[self.mask]
        # inner contour
        inn = np.nonzero(self.inner)[0]
        bi = b0[inn] + 1 + self.nv[inn]
        v["type"][bi] = capi.RECORD_STYLE_CHANGE
        v["has_move_to"][bi] = 1
        v["move_to_x"][bi] = self.ivx[inn, 0]
        v["move_to_y"][bi] = self.ivy[inn, 0]
        v["has_left_fill"][bi] = 1
        v["left_fill"][bi] = 2
        v["has_right_fill"][bi] = 1
        v["right_fill"][bi] = 1
        m2 = self.mask[inn]
        dest2 = (bi[:, None] + 1 + k)[m2]
        v["type"][dest2] = capi.RECORD_EDGE
        v["delta_x"][dest2] = self.idx_[inn][m2]
        v["delta_y"][dest2] = self.idy_[inn][m2]
        v["has_control_delta"][dest2] = self.curved[inn][m2]
        v["control_delta_x"][dest2] = self.icdx[inn][m2]
        v["control_delta_y"][dest2] = self.icdy[inn][m2]

        ids = np.zeros(n, dtype=np.uint32)
        lib = capi.load()
        rec_ptr = C.cast(recs, C.POINTER(capi.ShapeRecord))
        tag = capi.DefineShape()
        out = C.c_uint32()
        for i in range(n):
            keep = []
            nf = 2 if self.inner[i] else 1
            fills = (capi.FillStyle * nf)()
            fills[0] = _fill(self._fill_ast(i, False), keep)
            if nf == 2:
                fills[1] = _fill(self._fill_ast(i, True), keep)
            tag.id = i & 0xFFFF
            tag.initial_styles.n_fill = nf
            tag.initial_styles.fill = C.cast(fills, C.POINTER(capi.FillStyle))
            tag.initial_styles.n_line = 0
            if self.stroked[i]:
                line = (capi.LineStyle * 1)()
                la = self._line_ast(i)
                line[0].width = la["width"]
                line[0].morph_width = la["width"]
                line[0].fill = _fill(la["fill"], keep)
                keep.append(line)
                tag.initial_styles.n_line = 1
                tag.initial_styles.line = C.cast(line, C.POINTER(capi.LineStyle))
            tag.n_records = int(per[i])
            tag.records = C.cast(C.addressof(rec_ptr.contents) + int(base[i]) * C.sizeof(R), C.POINTER(capi.ShapeRecord))
            rc = lib.swfr_register_shape(renderer._h, C.byref(tag), C.byref(out))
            if rc != 0:
                raise RuntimeError("swfr_register_shape failed: %d" % rc)
            ids[i] = out.value
        return ids
