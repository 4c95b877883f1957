#!/bin/bash
# Runs on the GPU box (under gpurun): GPU tests, the bench, then the ncu launch list and one --set full capture
# of every kernel of one render pass.  Usage: tools/gpu_profile.sh <tag> [skip_tests] [skip_full_capture]
# Outputs under gpurun_out/: tests_<tag>.log bench_<tag>.log launches_<tag>.csv prof_<tag>.ncu-rep
set -u
TAG=${1:-x}
SKIP_TESTS=${2:-0}
SKIP_FULL=${3:-0}
mkdir -p gpurun_out
if [ "$SKIP_TESTS" != "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$TAG.log 2>&1
  echo "tests rc=$?" >> gpurun_out/tests_$TAG.log
  tail -3 gpurun_out/tests_$TAG.log
fi
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2>&1
echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_$TAG.log
# small profiling workload: 16 frames, one pass; kernels of one render = launches reported by the bench line
PCMD="python bench.py --frames 16 --frames-per-pass 16 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $PCMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
L=$(python -c "import json,sys; print(json.loads(open('gpurun_out/plain_$TAG.log').read().strip().splitlines()[-1])['gpu_launches'])")
echo "launches per render: $L"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * L)) -c $L --csv \
  --log-file gpurun_out/launches_$TAG.csv $PCMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
[ "$SKIP_FULL" = "1" ] && exit 0
timeout 900 ncu --set full --clock-control none --import-source on -s $((3 * L)) -c $L \
  -o gpurun_out/prof_$TAG -f $PCMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
