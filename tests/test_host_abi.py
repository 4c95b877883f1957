"""CPU tests of the product's host side: the C ABI loads and exports what include/swfr.h declares, and the
library's shape compiler / stroker / bitmap decoder (all host code) match the reference goldens and the oracle.
No compute call is made here (no GPU in this container)."""
import ctypes
import os
import re

import numpy as np
import pytest

import corpus
from oracle import compile_shape as cs
from oracle import decode_bitmap, raster

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_library):
    header = open(os.path.join(ROOT, "include", "swfr.h")).read()
    declared = set(re.findall(r"^(?:int|void|uint32_t|const char \*)\s*(swfr_[a-z0-9_]+)\s*\(", header, re.M))
    assert len(declared) >= 25
    raw = ctypes.CDLL(os.path.join(ROOT, "swf_renderer_b200", "libswfr_b200.so"))
    missing = [s for s in sorted(declared) if not hasattr(raw, s)]
    assert not missing, missing
    from swf_renderer_b200 import capi

    assert set(capi.PROTOTYPES) == declared
    assert built_library.swfr_abi_version() == 3


def test_no_cuda_device_is_a_loud_error_not_a_fallback(built_library):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import swf_renderer_b200 as sw

    with pytest.raises(sw.SwfrError) as e:
        sw.HeadlessRenderer(64, 64)
    assert e.value.status == -6  # SWFR_ERR_CUDA


def _oracle_commands(comp, morph):
    out = []
    for p in comp["paths"]:
        for c in p["commands"]:
            if morph:
                g = lambda k, i: float(c[k][i]) if k in c else 0.0
                if c["type"] == 2:
                    out.append([2, g("x", 0), g("y", 0), 0, 0, g("x", 1), g("y", 1), 0, 0])
                else:
                    out.append(
                        [c["type"], g("endX", 0), g("endY", 0), g("controlX", 0), g("controlY", 0)]
                        + [g("endX", 1), g("endY", 1), g("controlX", 1), g("controlY", 1)]
                    )
            else:
                g = lambda k: float(c.get(k, 0))
                if c["type"] == 2:
                    out.append([2, g("x"), g("y"), 0, 0, g("x"), g("y"), 0, 0])
                else:
                    row = [c["type"], g("endX"), g("endY"), g("controlX"), g("controlY")]
                    out.append(row + row[1:])
    return np.array(out, dtype=np.float64)


def _commands_to_golden_json(cmds, info, comp_styles, morph):
    """Rebuild the reference JSON from the product's command dump + the styles (styles come from the AST)."""
    paths = []
    at = 0
    for (n, has_fill, has_line), style in zip(info, comp_styles):
        commands = []
        for row in cmds[at : at + n]:
            t = int(row[0])
            num = lambda v: int(v) if float(v).is_integer() else float(v)
            if morph:
                pair = lambda i: [num(row[1 + i]), num(row[5 + i])]
                if t == 2:
                    commands.append({"type": 2, "x": pair(0), "y": pair(1)})
                elif t == 0:
                    commands.append({"type": 0, "endX": pair(0), "endY": pair(1)})
                else:
                    commands.append({"type": 1, "controlX": pair(2), "controlY": pair(3), "endX": pair(0), "endY": pair(1)})
            else:
                if t == 2:
                    commands.append({"type": 2, "x": num(row[1]), "y": num(row[2])})
                elif t == 0:
                    commands.append({"type": 0, "endX": num(row[1]), "endY": num(row[2])})
                else:
                    commands.append(
                        {"type": 1, "controlX": num(row[3]), "controlY": num(row[4]), "endX": num(row[1]), "endY": num(row[2])}
                    )
        at += n
        paths.append({"commands": commands, **style})
    return cs.to_golden_json({"paths": paths})


@pytest.mark.parametrize("sample", [s for s, _ in corpus.SHAPE_SAMPLES])
def test_library_compile_matches_reference_golden(built_library, sample):
    import swf_renderer_b200 as sw

    tag = corpus.load_ast(sample)
    cmds, info, segs = sw.compile_tag(tag)
    comp = cs.compile_shape(tag)
    np.testing.assert_array_equal(cmds, _oracle_commands(comp, False))
    styles = [{k: v for k, v in p.items() if k != "commands"} for p in comp["paths"]]
    assert [bool(i[1]) for i in info] == ["fill" in s for s in styles]
    # byte-exact against the reference's own golden (decode-shape.spec.ts:18-22)
    assert _commands_to_golden_json(cmds, info, styles, False) == corpus.read_text(sample, "shape.ts.json")


def test_library_morph_compile_matches_reference_golden(built_library):
    import swf_renderer_b200 as sw

    tag = corpus.load_ast(corpus.MORPH_SAMPLE)
    cmds, info, segs = sw.compile_tag(tag, morph=True)
    comp = cs.compile_morph_shape(tag)
    np.testing.assert_array_equal(cmds, _oracle_commands(comp, True))
    styles = [{k: v for k, v in p.items() if k != "commands"} for p in comp["paths"]]
    assert _commands_to_golden_json(cmds, info, styles, True) == corpus.read_text(corpus.MORPH_SAMPLE, "shape.ts.json")
    assert (segs[:, 8:] != segs[:, 2:8]).any()  # end state differs from start state
    assert 1867.5 in segs  # synthesised control point delta/2 (decode-swf-morph-shape.ts:341-346)


@pytest.mark.parametrize("sample", [s for s, _ in corpus.SHAPE_SAMPLES])
def test_library_segments_match_oracle(built_library, sample):
    """Implicit close + stroke expansion: the library's device segments equal the oracle's, bit for bit."""
    import swf_renderer_b200 as sw

    tag = corpus.load_ast(sample)
    _, _, segs = sw.compile_tag(tag)
    bm = {}
    if sample.startswith("textured"):
        bt = corpus.load_bitmap_ast("bitmap/homestuck-beta-3")
        bm[bt["id"]] = decode_bitmap.define_bitmap_rgba(bt)
    b = raster._Builder(bm)
    raster.add_shape_def(b, cs.compile_shape(tag))
    osegs = np.array([[k, path] + list(s6) + list(e6) for (s6, e6, k, path) in b.segs], dtype=np.float64)
    np.testing.assert_array_equal(segs, osegs)


def test_library_bitmap_decoder_matches_pam(built_library):
    import swf_renderer_b200 as sw

    tag = corpus.load_bitmap_ast("bitmap/homestuck-beta-3")
    rgba = sw.decode_x_swf_bmp(bytes.fromhex(tag["data"]))
    with open(corpus.CORPUS + "/bitmap/homestuck-beta-3.pam", "rb") as f:
        assert decode_bitmap.to_pam(rgba) == f.read()
    bad = bytearray(bytes.fromhex(tag["data"]))
    bad[0] = 5
    with pytest.raises(sw.SwfrError):
        sw.decode_x_swf_bmp(bytes(bad))


def test_library_compile_errors(built_library):
    import swf_renderer_b200 as sw

    tag = corpus.load_ast("flat-shapes/triangle")
    bad = {**tag, "shape": {**tag["shape"], "records": [dict(tag["shape"]["records"][0], left_fill=9)]}}
    with pytest.raises(sw.SwfrError) as e:
        sw.compile_tag(bad)
    assert e.value.status == -3  # Invalid fill ID
    morph = corpus.load_ast(corpus.MORPH_SAMPLE)
    recs = [dict(r) for r in morph["shape"]["records"]]
    del recs[0]["morph_move_to"]
    with pytest.raises(sw.SwfrError) as e:
        sw.compile_tag({**morph, "shape": {**morph["shape"], "records": recs}}, morph=True)
    assert e.value.status == -8  # Expected morphMoveTo to be defined
