#!/bin/bash
# Regenerates the committed ncu evidence from a capture brought back by tools/gpu_exp.sh <tag> ... full:
#   tools/make_profiles.sh <tag> <round-tag>     e.g.  tools/make_profiles.sh r2l r02
# Writes profiles/<round-tag>_kernels.md, <round-tag>_launches.csv, <round-tag>_fine_lines.md, <round-tag>_chain_lines.md
# and profiles/ncu_summary.json (with the hash of the kernel sources the capture was taken on - run this before
# touching csrc/kernels.cu or kernels.h again).
set -eu
TAG=$1
RT=$2
REP=gpurun_out/prof_$TAG.ncu-rep
python tools/ncu_summary.py $REP gpurun_out/launches_$TAG.csv $RT > /dev/null
{
  echo "# k_fine, per source line ($RT, \`$(basename $REP)\`, csrc/kernels.cu at the commit of this file)"
  echo
  echo "Share of the kernel's warp instructions, average active lanes per instruction and share of the stall samples, for"
  echo "every line with at least 0.5 % of either; both k_fine launches of the captured pass."
  echo
  echo '```'
  python tools/ncu_lines.py $REP k_fine 0.5
  echo '```'
} > profiles/${RT}_fine_lines.md
{
  echo "# chain kernels, per source line ($RT, \`$(basename $REP)\`)"
  echo
  for K in k_cover k_bin k_flatten_emit k_path_alive k_path_setup; do
    echo "## $K"
    echo '```'
    python tools/ncu_lines.py $REP $K 2.0
    echo '```'
  done
} > profiles/${RT}_chain_lines.md
ls -la profiles/${RT}_*
