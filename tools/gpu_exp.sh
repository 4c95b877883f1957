#!/bin/bash
# Runs on the GPU box (under gpurun).  Usage: tools/gpu_exp.sh <tag> <tests: 0|1|"-k expr"> <ncu: 0|list|fine|full> [variant ...]
# A variant is NAME:ENV=VAL,ENV=VAL (bench.py is run once per variant with those environment variables, short form).
# Outputs under gpurun_out/: tests_<tag>.log bench_<tag>[_<variant>].log launches_<tag>.csv prof_<tag>*.ncu-rep
set -u
TAG=${1:-x}
TESTS=${2:-1}
NCU=${3:-0}
shift 3 || true
mkdir -p gpurun_out
if [ "$TESTS" != "0" ]; then
  if [ "$TESTS" = "1" ]; then
    timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$TAG.log 2>&1
  else
    timeout 1500 python -m pytest tests -m gpu -x -q $TESTS > gpurun_out/tests_$TAG.log 2>&1
  fi
  echo "tests rc=$?" >> gpurun_out/tests_$TAG.log
  tail -4 gpurun_out/tests_$TAG.log
fi
SHORT="--steps 40 --warmup 3 --no-cpu-baseline --uhd-frames 0 --quick"
for V in "$@"; do
  NAME=${V%%:*}
  ENVS=${V#*:}
  [ "$ENVS" = "$V" ] && ENVS=""
  ( for kv in ${ENVS//,/ }; do export "$kv"; done; timeout 600 python bench.py $SHORT > gpurun_out/bench_${TAG}_$NAME.log 2>&1 )
  echo "variant $NAME rc=$?"
  python - gpurun_out/bench_${TAG}_$NAME.log <<'PY'
import json, sys
for l in open(sys.argv[1]):
    l = l.strip()
    if l.startswith("{"):
        d = json.loads(l)
        r = d.get("roofline", {})
        print("  ms/step %.3f  value %.0f  e2e %.0f (%.2f ms)  fine launch %.4f ms  stages %s  launches %s" % (
            d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], r.get("launch_ms", 0),
            {k: round(v, 3) for k, v in r.get("stage_ms_per_step", {}).items()}, d.get("gpu_launches")))
PY
done
[ "$NCU" = "0" ] && exit 0
# one pass of the default size (32 frames): every kernel of a render once, k_fine twice (slices of 16 frames)
PCMD="python bench.py --frames 32 --steps 1 --warmup 3 --no-cpu-baseline --uhd-frames 0 --quick"
timeout 300 $PCMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
L=$(python -c "import json,sys; print(json.loads(open('gpurun_out/plain_$TAG.log').read().strip().splitlines()[-1])['launches_per_render'])")
echo "launches per render: $L"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * L)) -c $L --csv \
  --log-file gpurun_out/launches_$TAG.csv $PCMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
if [ "$NCU" = "fine" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fine -s 3 -c 1 \
    -o gpurun_out/prof_${TAG}_fine -f $PCMD > gpurun_out/ncu_fine_$TAG.log 2>&1
  echo "ncu fine rc=$?"
elif [ "$NCU" = "full" ]; then
  # the hash of the kernel sources this capture is taken on (bench.py quotes the figures only for matching sources)
  python -c "import bench; print(bench.kernels_sha())" > gpurun_out/prof_$TAG.sha
  timeout 1200 ncu --set full --clock-control none --import-source on -s $((3 * L)) -c $L \
    -o gpurun_out/prof_$TAG -f $PCMD > gpurun_out/ncu_full_$TAG.log 2>&1
  echo "ncu full rc=$?"
fi
