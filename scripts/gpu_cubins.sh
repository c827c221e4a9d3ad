#!/bin/bash
# time the pp2_kernel instantiations found in every cubin under kb_variants/ (tools/sass_patch.py experiments) against
# each other; optional: one ncu --set full capture of the velocity-only product kernel
set -u
OUT=gpurun_out; mkdir -p $OUT
LIST=$(ls kb_variants/*.cubin | tr '\n' ':')
KBENCH_CUBIN=$LIST KBENCH_CUBIN_ONLY=1 timeout 600 omega3d_b200/csrc/microbench/kbench ${1:-262144} 5 2>&1 | tee $OUT/cubins.txt
if [ "${2:-}" = "ncu-vel" ]; then
  KBENCH_CUBIN=kb_variants/product_tuned.cubin KBENCH_CUBIN_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on \
      -k regex:pp2_kernel -s 7 -c 1 -f -o $OUT/pp2_vel_full omega3d_b200/csrc/microbench/kbench 262144 5 > $OUT/ncu_vel.log 2>&1
  ncu -i $OUT/pp2_vel_full.ncu-rep --page raw --csv > $OUT/pp2_vel_full_raw.csv 2>/dev/null
  tail -3 $OUT/ncu_vel.log
fi
