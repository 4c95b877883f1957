#!/bin/bash
# Builds libswfr_b200.so with AddressSanitizer + UBSan (host code; -Wall -Wextra) into a scratch copy of the tree and
# runs the CPU tests (host compiler, stroker, style / bitmap decoders, display-tree flattening, hostile inputs)
# against it.  Needs no GPU.  Prints the sanitizer findings, exit code 1 if there is any.
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
WORK=${1:-$ROOT/gpurun_out/asan}
rm -rf "$WORK" && mkdir -p "$WORK"
(cd "$ROOT" && tar --exclude=.git --exclude=gpurun_out --exclude=__pycache__ --exclude='*.so' -cf - .) | tar -xf - -C "$WORK"
cp "$ROOT"/oracle/*.so "$WORK/oracle/" 2>/dev/null
(cd "$WORK/swf_renderer_b200/csrc" && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -lineinfo -fmad=false \
  -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-Wall,-Wextra,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer \
  -shared -o ../libswfr_b200.so kernels.cu renderer.cu compile.cpp stroker.cpp styles.cpp display.cpp -lz -ldl) || exit 2
D=$(dirname "$(gcc -print-file-name=libasan.so)")
cd "$WORK" && ASAN_OPTIONS=detect_leaks=0:halt_on_error=0:protect_shadow_gap=0 UBSAN_OPTIONS=print_stacktrace=1 \
  LD_PRELOAD="$D/libasan.so $D/libubsan.so" python -m pytest tests -m "not gpu" -q -s \
  --deselect tests/test_sharding_gloo.py --deselect tests/test_bench_reference_arm.py --deselect tests/test_build_resources.py \
  > "$WORK/asan_run.log" 2>&1
tail -2 "$WORK/asan_run.log"
N=$(grep -c "runtime error\|AddressSanitizer" "$WORK/asan_run.log")
echo "sanitizer findings: $N"
grep "runtime error\|AddressSanitizer" "$WORK/asan_run.log" | sort | uniq -c
[ "$N" = "0" ]
