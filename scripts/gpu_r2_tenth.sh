#!/bin/bash
# round 2, tenth GPU call: whole-blocks-first partition (CTAs in step on the source stream) against the pure stream-K partition,
# each with and without the per-tile barrier, and with the CTAs started a fraction of a tile time apart. Same box, 256 K / 1 M / 4 M x 512 K targets.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2k_partition.txt; : > $O
K=omega3d_b200/csrc/microbench
for n in 262144 1048576; do
  echo "== N=$n" >> $O
  KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=kb_variants/hyb_nobar.cubin:kb_variants/hyb_bar.cubin:kb_variants/hyb_nobar_skew100.cubin:kb_variants/hyb_nobar_skew400.cubin timeout 300 $K/kbench $n 3 2>&1 | grep "velgrad" >> $O
  KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=kb_variants/pure/pure_nobar.cubin:kb_variants/pure/pure_bar.cubin timeout 300 kb_variants/pure/kbench $n 3 2>&1 | grep "velgrad" >> $O
  KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=kb_variants/hyb_nobar.cubin timeout 300 $K/kbench $n 3 2>&1 | grep "velgrad" >> $O
done
cat $O
