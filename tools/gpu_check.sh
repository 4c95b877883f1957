#!/bin/bash
# GPU tests + bench, no profiler.  Usage: tools/gpu_check.sh <tag> [pytest -k expression]
TAG=${1:-x}
KEXPR=${2:-}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/tests_$TAG.log 2>&1
else
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$TAG.log 2>&1
fi
echo "tests rc=$?" >> gpurun_out/tests_$TAG.log
tail -4 gpurun_out/tests_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2>&1
echo "bench rc=$?"
tail -c 2600 gpurun_out/bench_$TAG.log
