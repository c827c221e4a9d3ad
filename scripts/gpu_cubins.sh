#!/bin/bash
# gpurun -- bash scripts/gpu_cubins.sh <dir>[:kind[:nouniform]] ...   - time the pp2_kernel of every cubin under <dir> (a population
# of tools/tune_gpu.py) with kbench, twice; the ranking is read by `tools/tune_gpu.py pick`.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
for d in "$@"; do
  # <dir>[:kind[:nouniform]]  kind = velgrad (default) | vel; nouniform: force the per-particle-radius path
  kind=velgrad; nouni=""
  case $d in *:*) IFS=: read d kind nouni <<< "$d";; esac
  LIST=$(ls $d/*.cubin | tr '\n' ':')
  for rep in 1 2; do
    if [ -n "$nouni" ]; then export KBENCH_NO_UNIFORM=1; else unset KBENCH_NO_UNIFORM; fi
    KBENCH_CUBIN=$LIST KBENCH_CUBIN_ONLY=1 timeout 900 omega3d_b200/csrc/microbench/kbench 262144 3 2>&1 | grep " $kind "
  done > $OUT/cubins_$(basename $d).txt
  sort -k12 -n $OUT/cubins_$(basename $d).txt | head -5
done
