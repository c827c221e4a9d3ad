#!/bin/bash
# gpurun -- bash scripts/gpu_cubins.sh <dir>[:kind[:nouniform[:core]]] ...   (core = 1 | 2 | 3: ppc_kernel cubins of that core function)   - time the pp2_kernel of every cubin under <dir> (a population
# of tools/tune_gpu.py) with kbench, twice; the ranking is read by `tools/tune_gpu.py pick`.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
for d in "$@"; do
  # <dir>[:kind[:nouniform[:core]]]  kind = velgrad (default) | vel; nouniform: force the per-particle-radius path (empty = off)
  kind=velgrad; nouni=""; core=""
  case $d in *:*) IFS=: read d kind nouni core <<< "$d";; esac
  LIST=$(ls $d/*.cubin | tr '\n' ':')
  for rep in 1 2; do
    if [ -n "$nouni" ]; then export KBENCH_NO_UNIFORM=1; else unset KBENCH_NO_UNIFORM; fi
    if [ -n "$core" ]; then export KBENCH_CORE=$core; else unset KBENCH_CORE; fi
    KBENCH_CUBIN=$LIST KBENCH_CUBIN_ONLY=1 timeout 900 omega3d_b200/csrc/microbench/kbench 262144 3 2>&1 | grep " $kind "
  done > $OUT/cubins_$(basename $d).txt
  sort -k12 -n $OUT/cubins_$(basename $d).txt | head -5
done
