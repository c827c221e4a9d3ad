#!/bin/bash
# round 2, fifth GPU call: where do the 3 % between the round-1 one-block-per-CTA kernel and the persistent stream-K kernel go?
# Same box, same sizes: round-1 kernel (its own harness), current kernel, start-staggered CTAs, 256-source tiles; ncu --set full
# with per-instruction sampling of the current kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2e_kbench.txt; : > $O
K=omega3d_b200/csrc/microbench
for n in 262144 1048576; do
  echo "== r1 kernel N=$n" >> $O
  KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=kb_variants/r1/tuned.cubin timeout 300 kb_variants/r1/kbench $n 3 2>&1 | grep velgrad >> $O
  echo "== current N=$n" >> $O
  KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=kb_variants/cur.cubin:kb_variants/stag40k.cubin:kb_variants/stag13k.cubin:kb_variants/cur.cubin timeout 300 $K/kbench $n 3 2>&1 | grep velgrad >> $O
done
echo "== 256-source tiles N=262144" >> $O
KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=kb_variants/t256/t256.cubin timeout 300 kb_variants/t256/kbench 262144 3 2>&1 | grep velgrad >> $O
cat $O
KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=kb_variants/cur.cubin timeout 600 ncu --set full --clock-control none --import-source on -k regex:pp2_kernel -s 1 -c 1 -f \
   -o gpurun_out/r2e_pp2_cur $K/kbench 262144 1 > gpurun_out/r2e_ncu.log 2>&1
ncu -i gpurun_out/r2e_pp2_cur.ncu-rep --page raw --csv > gpurun_out/r2e_pp2_cur_raw.csv 2>/dev/null
ls -la gpurun_out/r2e_*
