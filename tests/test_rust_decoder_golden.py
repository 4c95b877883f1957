"""The Rust crate's decoder goldens (tests/flat-shapes/*/shape.rs.log, rs/src/lib.rs:38-69): the oracle's restatement
of rs/src/decoder/shape_decoder.rs reproduces them byte for byte, and the product's compiler (TypeScript semantics,
curves kept) yields the same chains once its curves are read as straight segments."""
import numpy as np
import pytest

import corpus
from oracle import compile_shape as cs

FLAT = [s for s, _ in corpus.SHAPE_SAMPLES if s.startswith("flat-shapes/")]


@pytest.mark.parametrize("sample", FLAT)
def test_rust_decoder_golden(sample):
    tag = corpus.load_ast(sample)
    assert cs.rust_debug(cs.rust_decode_shape(tag)) == corpus.read_text(sample, "shape.rs.log")


@pytest.mark.parametrize("sample", FLAT)
def test_library_compiler_agrees_with_rust_decoder(built_library, sample):
    """swfr_compile_debug (no GPU needed): per path the same MoveTo / LineTo chain as the Rust decoder, with CurveTo
    end points where the Rust decoder draws a straight line."""
    from swf_renderer_b200 import compile_tag

    tag = corpus.load_ast(sample)
    rust = cs.rust_decode_shape(tag)
    cmds, info, _ = compile_tag(tag)
    assert len(info) == len(rust)
    at = 0
    for k, rp in enumerate(rust):
        n = int(info[k][0])
        rows = cmds[at:at + n]
        at += n
        verbs = ["MoveTo" if int(r[0]) == cs.MOVE_TO else "LineTo" for r in rows]
        points = [(float(r[1]), float(r[2])) for r in rows]
        assert verbs == rp["verbs"]
        assert points == [(float(x), float(y)) for x, y in rp["points"]]
        assert bool(info[k][1]) == (rp["fill"] is not None) and bool(info[k][2]) == (rp["line"] is not None)
