#!/bin/bash
# On the GPU box (under gpurun): the end-to-end leg of the bench in its short form, then the GPU tests that exercise the
# host side of the library (renders in flight, read-backs, recovery, gather, display tree).  tools/e2e_check.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 100 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --uhd-frames 0 --quick > gpurun_out/e2e_$TAG.log 2> gpurun_out/e2e_$TAG.err
echo "bench rc=$?"
python - gpurun_out/e2e_$TAG.log <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print("device ms/step %.3f  e2e ms/step %.3f  parity %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity"]))
PY
timeout ${2:-300} python -m pytest tests/test_gpu_parity.py tests/test_display_tree.py tests/test_color_transform.py \
  tests/test_hostile_inputs.py tests/test_gpu_synthetic.py -m gpu -x -v --durations=12 > gpurun_out/tests_$TAG.log 2>&1
echo "tests rc=$?"
tail -22 gpurun_out/tests_$TAG.log
