#!/bin/bash
# Regenerates tests/golden/corpus from the reference checkout (run in the build container, where /root/reference
# exists; the GPU box has no reference tree, which is why the fixtures are committed).  Only test DATA is copied -
# inputs (ast.json) and the reference's own expected outputs (shape.ts.json, shape.rs.log, *.png, *.pam) - no source.
set -eu
REF=${1:-/root/reference/tests}
DST=$(cd "$(dirname "$0")" && pwd)/corpus
for d in flat-shapes/homestuck-beta-1 flat-shapes/squares flat-shapes/triangle; do
  mkdir -p "$DST/$d"
  cp "$REF/$d/ast.json" "$REF/$d/shape.ts.json" "$REF/$d/shape.rs.log" "$REF/$d/shape.png" "$DST/$d/"
done
d=textured-shapes/homestuck-beta-4
mkdir -p "$DST/$d" && cp "$REF/$d/ast.json" "$REF/$d/shape.ts.json" "$REF/$d/shape.png" "$DST/$d/"
d=flat-morph-shapes/homestuck-beta-29
mkdir -p "$DST/$d" && cp "$REF/$d/ast.json" "$REF/$d/shape.ts.json" "$REF/$d/0.png" "$REF/$d/32768.png" "$REF/$d/65536.png" "$DST/$d/"
mkdir -p "$DST/bitmap" && cp "$REF/bitmap/homestuck-beta-3.ast.json" "$REF/bitmap/homestuck-beta-3.pam" "$DST/bitmap/"
echo "copied $(find "$DST" -type f | wc -l) files"
