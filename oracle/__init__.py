"""CPU oracle for the SWF shape -> pixels path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import, link or execute it, and only as the checker
or the reported CPU baseline.  The product (``swf_renderer_b200`` + ``libswfr_b200.so``)
never routes through it and fails loudly when its CUDA library is missing.

What it restates (reference = open-flash/swf-renderer, TypeScript renderer):

* ``compile_shape.py``   - ``ts/src/lib/shape/decode-swf-shape.ts`` and
  ``decode-swf-morph-shape.ts`` (records -> ordered style paths).  Pinned: byte-exact
  against the five ``shape.ts.json`` goldens.
* ``decode_bitmap.py``   - ``ts/src/lib/decode-x-swf-bmp.ts``.  Pinned: byte-exact
  against ``tests/bitmap/homestuck-beta-3.pam``.
* ``stroker.py``         - ``ctx.stroke()`` semantics of ``canvas-renderer.ts:252-266,339-349``
  (stroke-to-fill expansion in user space).
* ``raster.c``           - the draw loop of ``canvas-renderer.ts:61-350`` plus the part of
  node-canvas/Cairo/pixman (npm ``canvas@2.6.1``, NOT vendored in the reference tree) that
  the draw loop calls into: flatten, non-zero AA fill, pattern/gradient sampling and
  premultiplied 8-bit OVER.  Cairo's exact scan converter is not reproduced; coverage is
  exact-area.  Pinned against the reference's seven PNG goldens within the tolerance
  north_star states (interior |d| <= 2/255, PSNR >= 40 dB, premultiplied compare).

Parity status: geometry / compile / bitmap decode / solid fills / non-repeating minified
bitmap fill / quadratic morph curves / butt+miter strokes are PINNED by reference goldens.
Gradients (all kinds), spread modes, linear-RGB interpolation, repeating bitmaps and
translucent fills have no reference fixture ("parity unpinned" for those features; the
oracle follows the Canvas/SWF semantics written in DESIGN.md).
"""
