#!/bin/bash
# time pp2_kernel<2,true,128> from every cubin under kb_variants/ (tools/sass_patch.py experiments) against each other
set -u
OUT=gpurun_out; mkdir -p $OUT
LIST=$(ls kb_variants/*.cubin | tr '\n' ':')
KBENCH_CUBIN=$LIST KBENCH_CUBIN_ONLY=1 timeout 600 omega3d_b200/csrc/microbench/kbench ${1:-262144} 5 2>&1 | tee $OUT/cubins.txt
