"""Hostile inputs to the host-only entry points (no GPU): corrupted x-swf-bmp payloads and mutated shape / morph-shape
tags must come back as a status code (or a valid result), never crash, hang or read out of bounds.  The same file is
what the AddressSanitizer / UBSan build of the library is run against ("Oracle and tests").

Reference behaviour being mapped: decode-x-swf-bmp.ts:9-41 throws on a wrong format id and lets zlib throw on a bad
stream; decode-swf-shape.ts:418-420 throws "Invalid fill ID"; decode-swf-morph-shape.ts:315-318 throws on a missing
morphMoveTo."""
import copy
import zlib

import numpy as np
import pytest

import corpus
from test_compile_fuzz import _random_tag


def _xswfbmp(width, height, colors, table, indices, fmt=3):
    """format id, width (LE16), height (LE16), colour count - 1, zlib(table RGB + padded indices)."""
    head = bytes([fmt]) + int(width).to_bytes(2, "little") + int(height).to_bytes(2, "little") + bytes([colors - 1])
    return head + zlib.compress(bytes(table) + bytes(indices))


def test_xswfbmp_decoder_survives_corrupted_payloads(built_library):
    import swf_renderer_b200 as sw

    rng = np.random.RandomState(7)
    w, h, colors = 5, 3, 4
    padded = (w + 3) & ~3
    table = rng.randint(0, 256, 3 * colors).tolist()
    idx = rng.randint(0, colors, padded * h).tolist()
    good = _xswfbmp(w, h, colors, table, idx)
    img = sw.decode_x_swf_bmp(good)
    assert img.shape == (h, w, 4) and (img[..., 3] == 255).all()
    # an index past the colour table is opaque black (decode-x-swf-bmp.ts:35-36)
    idx2 = list(idx)
    idx2[0] = 200
    assert sw.decode_x_swf_bmp(_xswfbmp(w, h, colors, table, idx2))[0, 0].tolist() == [0, 0, 0, 255]
    n_ok = n_err = 0
    cases = [good[:k] for k in range(len(good))]  # every truncation
    cases += [_xswfbmp(w, h, colors, table, idx[: len(idx) // 2])]  # inflated data shorter than the header promises
    cases += [_xswfbmp(w, h, colors, table[:3], idx)]
    cases += [_xswfbmp(0, 0, 1, [1, 2, 3], [])]
    cases += [_xswfbmp(65535, 65535, 256, [0] * 768, [0] * 16)]  # a header that promises 4 Gpixel
    for k in range(300):  # random byte flips, header included
        b = bytearray(good)
        for _ in range(rng.randint(1, 4)):
            b[rng.randint(0, len(b))] = rng.randint(0, 256)
        cases.append(bytes(b))
    for data in cases:
        try:
            out = sw.decode_x_swf_bmp(data)
            assert out.ndim == 3 and out.shape[2] == 4
            n_ok += 1
        except sw.SwfrError as e:
            assert e.status < 0
            n_err += 1
    assert n_err >= len(good)  # every truncation is refused
    assert n_ok + n_err == len(cases)


def _mutate(tag, rng, morph):
    """One random structural mutation of a valid tag (still JSON-shaped, so it goes through the same conversion)."""
    t = copy.deepcopy(tag)
    recs = t["shape"]["records"]
    kind = rng.randint(0, 7)
    if kind == 0 and recs:  # style index out of range
        sc = [r for r in recs if r["type"] == "style-change"]
        if sc:
            r = sc[rng.randint(0, len(sc))]
            r[["left_fill", "right_fill", "line_style"][rng.randint(0, 3)]] = int(rng.choice([7, 255, 65535, 2 ** 31 - 1]))
    elif kind == 1:  # no records at all
        t["shape"]["records"] = []
    elif kind == 2 and recs:  # only style changes / only edges
        keep = "edge" if rng.rand() < 0.5 else "style-change"
        t["shape"]["records"] = [r for r in recs if r["type"] == keep]
    elif kind == 3 and recs:  # extreme coordinates
        for r in recs:
            if r["type"] == "edge" and rng.rand() < 0.3:
                r["delta"] = {"x": int(rng.choice([-2 ** 31, 2 ** 31 - 1, 0])), "y": int(rng.choice([-2 ** 31, 2 ** 31 - 1, 0]))}
    elif kind == 4:  # empty style tables while records still refer to them
        t["shape"]["initial_styles"] = {"fill": [], "line": []}
    elif kind == 5 and recs:  # records shuffled
        rng.shuffle(recs)
    elif kind == 6 and morph and recs:  # morph bookkeeping removed
        for r in recs:
            if rng.rand() < 0.5:
                for k in ("morph_move_to", "morph_delta", "morph_control_delta"):
                    r.pop(k, None)
    return t


@pytest.mark.parametrize("morph", [False, True])
def test_mutated_tags_compile_or_fail_with_a_status(built_library, morph):
    import swf_renderer_b200 as sw

    rng = np.random.RandomState(11 + int(morph))
    n_ok = n_err = 0
    for k in range(150):
        tag = _mutate(_random_tag(5000 + k, morph), rng, morph)
        try:
            cmds, info, segs = sw.compile_tag(tag, morph=morph)
            assert np.isfinite(cmds).all()
            assert info.shape[1] == 3 and (info[:, 0] >= 0).all()
            n_ok += 1
        except sw.SwfrError as e:
            assert e.status < 0, e
            n_err += 1
    assert n_ok and n_err  # both outcomes occur; neither is a crash


def test_display_tree_depth_is_bounded_not_a_stack_overflow(built_library):
    """A container chain far deeper than any real display list: flattened or refused with a status, but the process
    survives (the reference recurses without a limit, canvas-renderer.ts:131-145)."""
    import ctypes as C

    from swf_renderer_b200 import capi
    from swf_renderer_b200.display import flatten_stage
    from swf_renderer_b200.renderer import SwfrError

    for depth in (64, 5000):
        keep = []
        leaf = (capi.DisplayObject * 1)()
        leaf[0].type = capi.DISPLAY_SHAPE
        leaf[0].id = 1
        keep.append(leaf)
        node = leaf
        for _ in range(depth):
            box = (capi.DisplayObject * 1)()
            box[0].type = capi.DISPLAY_CONTAINER
            box[0].n_children = 1
            box[0].children = C.cast(node, C.POINTER(capi.DisplayObject))
            keep.append(box)
            node = box
        st = capi.DisplayStage()
        st.width, st.height = 64, 64
        st.n_children = 1
        st.children = C.cast(node, C.POINTER(capi.DisplayObject))
        try:
            prims, n = flatten_stage(st)
            assert n == 1 and prims[0].id == 1
        except SwfrError as e:
            assert depth > 64 and e.status < 0


def test_xswfbmp_header_cannot_make_the_decoder_allocate_gigabytes(built_library):
    """A 6-byte header promising 65535 x 65535 pixels over a 20-byte payload: refused as truncated without sizing any
    buffer from the header - checked in a child process whose address space is capped at 3 GB (the 4.3 GB
    zero-fill of round 1 aborted there with std::bad_alloc crossing the C ABI)."""
    import os
    import subprocess
    import sys

    if "libasan" in os.environ.get("LD_PRELOAD", ""):
        pytest.skip("AddressSanitizer reserves terabytes of address space: no RLIMIT_AS under tools/asan_check.sh")
    code = (
        "import resource, zlib, sys\n"
        "sys.path.insert(0, %r)\n"
        "import swf_renderer_b200 as sw\n"
        "resource.setrlimit(resource.RLIMIT_AS, (3 << 30, 3 << 30))\n"
        "data = bytes([3, 255, 255, 255, 255, 255]) + zlib.compress(b'\\0' * 10)\n"
        "try:\n"
        "    sw.decode_x_swf_bmp(data)\n"
        "    print('decoded')\n"
        "except sw.SwfrError as e:\n"
        "    print('status', e.status)\n"
    ) % corpus.os.path.dirname(corpus.HERE)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-500:]
    assert p.stdout.strip().startswith("status -"), p.stdout
