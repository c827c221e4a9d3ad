#!/bin/bash
# gpurun --timeout 1200 -- bash scripts/gpu_tests.sh : parity tests (all, not -x), kernel shoot-out, a bench line
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== kbench"; timeout 300 omega3d_b200/csrc/microbench/kbench 262144 3 2>&1 | grep packed | tee $OUT/kbench.txt
echo "== kbench general-radius path"; KBENCH_NO_UNIFORM=1 timeout 300 omega3d_b200/csrc/microbench/kbench 262144 3 2>&1 | grep packed | tee $OUT/kbench_nouniform.txt
echo "== bench 1M"; timeout 300 python bench.py --no-cpu 2>&1 | tail -1 | tee $OUT/bench_1m.json
