#!/bin/bash
# like gpu_cubins.sh, but also through the per-particle-radius paths (KBENCH_NO_UNIFORM)
set -u
OUT=gpurun_out; mkdir -p $OUT
LIST=$(ls kb_variants/*.cubin | tr '\n' ':')
echo "== uniform radii" | tee $OUT/cubins2.txt
KBENCH_CUBIN=$LIST KBENCH_CUBIN_ONLY=1 timeout 600 omega3d_b200/csrc/microbench/kbench ${1:-262144} 5 2>&1 | tee -a $OUT/cubins2.txt
echo "== per-particle radii path" | tee -a $OUT/cubins2.txt
KBENCH_NO_UNIFORM=1 KBENCH_CUBIN=$LIST KBENCH_CUBIN_ONLY=1 timeout 600 omega3d_b200/csrc/microbench/kbench ${1:-262144} 5 2>&1 | tee -a $OUT/cubins2.txt
