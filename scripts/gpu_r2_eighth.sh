#!/bin/bash
# round 2, eighth GPU call: per-tile barrier against "last warp out refills" (no CTA-wide barrier; warps may drift one tile apart):
# per-warp tile-completion skew (instrumented harness, linked kernels) and post-processed cubins timed at 256 K and 1 M.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2i_nobar.txt; : > $O
for h in kbench_endtime kbench_endtime_nobar; do
  echo "== $h N=262144" >> $O
  KBENCH_PRODUCT_ONLY=1 timeout 300 kb_variants/$h 262144 2 2>&1 | grep -v "B=128\|SM 0..3\|^ *$" | head -24 >> $O
done
K=omega3d_b200/csrc/microbench
for n in 262144 1048576; do
  echo "== cubins N=$n" >> $O
  KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=kb_variants/bar.cubin:kb_variants/nobar.cubin:kb_variants/bar.cubin:kb_variants/nobar.cubin timeout 300 $K/kbench $n 3 2>&1 | grep cubin >> $O
done
cat $O
