#!/bin/bash
# run every kernel-variant binary under kb_variants/ (built locally with different -D flags) and keep the product-shape lines
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/variants.txt
for b in kb_variants/*; do
  echo "== $b" | tee -a $OUT/variants.txt
  timeout 120 $b 262144 3 2>&1 | grep -E "packedtrue +T=2 B=128 split= 1|packedtrue +T=1 B=128|packedfalse +T=4 B=128|packedtrue +T=3" | tee -a $OUT/variants.txt
done
