"""Oracle stroke-to-fill expansion (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates what ``ctx.stroke()`` does for the reference's two call sites:
  ts/src/lib/renderers/canvas-renderer.ts:339-349  shapes: lineWidth = width (twips, user space), Canvas defaults
                                                   lineCap "butt", lineJoin "miter", miterLimit 10
  ts/src/lib/renderers/canvas-renderer.ts:252-266  morph shapes: lerped width, lineCap/lineJoin "round"
The reference never calls closePath(), so every sub-path is OPEN (caps at both ends, even when the last
point equals the first).  ``lineWidth = 0`` is ignored by Canvas, i.e. the previous width stays in force
(1 twip when nothing was set).  The pen is defined in user space (twips); the CTM is applied to the
outline afterwards, exactly as Canvas/Cairo do for a pen under a non-uniform transform.

Output: closed contours (one per sub-path) made of line and quadratic segments, to be filled non-zero.
Curves are split until each piece turns by <= ~15 degrees and are offset as quadratics, so the outline
stays resolution independent.  Only + - * / sqrt are used (no trig) so that the C++ product stroker can
reproduce the same doubles; coordinates are finally rounded to float32.
"""
from __future__ import annotations

import math

import numpy as np

MITER_LIMIT = 10.0
COS_SPLIT = 0.9659258262890683  # cos(15 deg): curve pieces turn by at most this
COS_ARC = 0.7071067811865476  # cos(45 deg): round joins/caps are built from arcs of at most 45 deg


def _unit(dx, dy):
    l = math.sqrt(dx * dx + dy * dy)
    if l == 0.0:
        return None
    return (dx / l, dy / l)


def _split_quad(p0, c, p1, out, depth=0):
    """Split a quadratic until the turning between end tangents is small.  Appends (p0, c, p1, t0, t1)."""
    t0 = _unit(c[0] - p0[0], c[1] - p0[1])
    t1 = _unit(p1[0] - c[0], p1[1] - c[1])
    if t0 is None and t1 is None:
        ch = _unit(p1[0] - p0[0], p1[1] - p0[1])
        if ch is not None:
            out.append((p0, None, p1, ch, ch))
        return
    if t0 is None:
        t0 = t1
    if t1 is None:
        t1 = t0
    dot = t0[0] * t1[0] + t0[1] * t1[1]
    if dot >= COS_SPLIT or depth >= 8:
        out.append((p0, c, p1, t0, t1))
        return
    a = ((p0[0] + c[0]) * 0.5, (p0[1] + c[1]) * 0.5)
    b = ((c[0] + p1[0]) * 0.5, (c[1] + p1[1]) * 0.5)
    m = ((a[0] + b[0]) * 0.5, (a[1] + b[1]) * 0.5)
    _split_quad(p0, a, m, out, depth + 1)
    _split_quad(m, b, p1, out, depth + 1)


def _arc(center, u, v, w, out, depth=0):
    """Arc of radius w around center from unit vector u to unit vector v (the short way), as quadratics."""
    dot = u[0] * v[0] + u[1] * v[1]
    if dot < COS_ARC and depth < 6:
        mx, my = u[0] + v[0], u[1] + v[1]
        mid = _unit(mx, my)
        if mid is None:  # half turn: pick the perpendicular of u (callers avoid this by splitting caps)
            mid = (-u[1], u[0])
        _arc(center, u, mid, w, out, depth + 1)
        _arc(center, mid, v, w, out, depth + 1)
        return
    k = w / (1.0 + dot)
    ctrl = (center[0] + (u[0] + v[0]) * k, center[1] + (u[1] + v[1]) * k)
    end = (center[0] + v[0] * w, center[1] + v[1] * w)
    out.append(("Q", ctrl, end))


def _pieces_of_subpath(cmds):
    """Turn one sub-path's commands into offsettable pieces (p0, c|None, p1, t0, t1)."""
    pieces = []
    cur = None
    for c in cmds:
        if c[0] == "M":
            cur = c[1]
        elif c[0] == "L":
            t = _unit(c[1][0] - cur[0], c[1][1] - cur[1])
            if t is not None:
                pieces.append((cur, None, c[1], t, t))
            cur = c[1]
        else:
            _split_quad(cur, c[1], c[2], pieces)
            cur = c[2]
    return pieces


def _offset_side(pieces, w, round_join, out):
    """Emit the left-offset (by +w along the left normal) outline of consecutive pieces, with joins."""
    first = True
    prev_t = None
    for (p0, c, p1, t0, t1) in pieces:
        n0 = (t0[1] * w, -t0[0] * w)
        n1 = (t1[1] * w, -t1[0] * w)
        start = (p0[0] + n0[0], p0[1] + n0[1])
        if first:
            out.append(("L", start))
            first = False
        else:
            # join between prev_t and t0 around p0
            cross = prev_t[0] * t0[1] - prev_t[1] * t0[0]
            dot = prev_t[0] * t0[0] + prev_t[1] * t0[1]
            if dot > 0.0 and abs(cross) < 1e-12:
                out.append(("L", start))  # smooth continuation (pieces of one curve, collinear lines)
            elif cross > 0.0 or (cross == 0.0 and dot <= 0.0):
                # this side is the OUTER side of the turn (normal (ty,-tx), y-down): build the join
                if round_join:
                    u = (prev_t[1], -prev_t[0])
                    v = (t0[1], -t0[0])
                    _arc(p0, u, v, w, out)
                else:
                    if MITER_LIMIT * MITER_LIMIT * (1.0 + dot) >= 2.0:
                        k = w / (1.0 + dot)
                        out.append(("L", (p0[0] + (prev_t[1] + t0[1]) * k, p0[1] + (-prev_t[0] - t0[0]) * k)))
                    out.append(("L", start))
            else:
                # inner side: go through the vertex so the outline keeps a consistent winding
                out.append(("L", p0))
                out.append(("L", start))
        end = (p1[0] + n1[0], p1[1] + n1[1])
        if c is None:
            out.append(("L", end))
        else:
            dotn = t0[0] * t1[0] + t0[1] * t1[1]
            k = w / (1.0 + dotn)
            ctrl = (c[0] + (t0[1] + t1[1]) * k, c[1] + (-t0[0] - t1[0]) * k)
            out.append(("Q", ctrl, end))
        prev_t = t1


def _reverse(pieces):
    rev = []
    for (p0, c, p1, t0, t1) in reversed(pieces):
        rev.append((p1, c, p0, (-t1[0], -t1[1]), (-t0[0], -t0[1])))
    return rev


def stroke_subpath(cmds, width, round_style):
    """One open sub-path -> one closed contour as a command list [("M",p), ("L",p) | ("Q",c,p) ...]."""
    pieces = _pieces_of_subpath(cmds)
    if not pieces:
        return []
    w = width * 0.5
    out = []
    _offset_side(pieces, w, round_style, out)
    # end cap
    t_end = pieces[-1][4]
    p_end = pieces[-1][2]
    if round_style:
        u = (t_end[1], -t_end[0])
        _arc(p_end, u, t_end, w, out)
        _arc(p_end, t_end, (-u[0], -u[1]), w, out)
    back = _reverse(pieces)
    tmp = []
    _offset_side(back, w, round_style, tmp)
    out.extend(tmp)  # first entry is ("L", start of the right side) == butt cap edge
    # start cap
    t_start = back[-1][4]
    p_start = back[-1][2]
    if round_style:
        u = (t_start[1], -t_start[0])
        _arc(p_start, u, t_start, w, out)
        _arc(p_start, t_start, (-u[0], -u[1]), w, out)
    first = out[0][1]
    contour = [("M", first)] + out[1:]
    contour.append(("L", first))
    return contour


def stroke_path(commands, width, round_style):
    """Compiled path commands [("M",(x,y)), ("L",(x,y)), ("Q",(cx,cy),(x,y))] -> list of contours."""
    subpaths = []
    cur = []
    for c in commands:
        if c[0] == "M":
            if len(cur) > 1:
                subpaths.append(cur)
            cur = [c]
        else:
            cur.append(c)
    if len(cur) > 1:
        subpaths.append(cur)
    contours = []
    for sp in subpaths:
        ct = stroke_subpath(sp, width, round_style)
        if ct:
            contours.append(ct)
    return contours


def contours_to_segments(contours):
    """Contours -> [(is_curve, x0,y0,cx,cy,x1,y1)] with coordinates rounded to float32."""
    f32 = lambda v: float(np.float32(v))
    segs = []
    for ct in contours:
        cur = None
        for c in ct:
            if c[0] == "M":
                cur = (f32(c[1][0]), f32(c[1][1]))
            elif c[0] == "L":
                p = (f32(c[1][0]), f32(c[1][1]))
                if p != cur:
                    segs.append((0, cur[0], cur[1], cur[0], cur[1], p[0], p[1]))
                cur = p
            else:
                cp = (f32(c[1][0]), f32(c[1][1]))
                p = (f32(c[2][0]), f32(c[2][1]))
                segs.append((1, cur[0], cur[1], cp[0], cp[1], p[0], p[1]))
                cur = p
    return segs
