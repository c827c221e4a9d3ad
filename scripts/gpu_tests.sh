#!/bin/bash
# gpurun --timeout 1200 -- bash scripts/gpu_tests.sh : parity tests (all, not -x), microbenchmarks, kernel shoot-out
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== issue model"; timeout 120 omega3d_b200/csrc/microbench/issue_model 2>&1 | tee $OUT/issue_model.txt
echo "== kbench"; timeout 300 omega3d_b200/csrc/microbench/kbench 262144 3 2>&1 | tee $OUT/kbench.txt
echo "== bench 256K"; timeout 300 python bench.py --particles 262144 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_256k.json
