#!/bin/bash
# round 2, sixth GPU call: do the three co-resident persistent CTAs of an SM progress equally? Per-CTA end times of the
# 128-thread x 3 per SM kernel, and the same walk as ONE 384-thread CTA per SM (a CTA-wide barrier per tile keeps its warps together).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 262144 1048576; do
KBENCH_PRODUCT_ONLY=1 timeout 300 kb_variants/kbench_endtime $n 3 2>&1
done | tee gpurun_out/r2f_endtime.txt
