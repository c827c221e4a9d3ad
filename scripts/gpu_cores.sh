#!/bin/bash
# Quick validation of the alternate core functions on a B200: the new GPU tests, the drop-in test, per-core timings.
# Usage: gpurun --timeout 420 -- bash scripts/gpu_cores.sh
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest cores + dropin"
timeout 300 python -m pytest tests/test_gpu_cores.py tests/test_gpu_dropin.py -q -m gpu -x 2>&1 | tail -15 | tee $OUT/pytest_cores.txt
echo "== bench cores 256K"
timeout 120 python tests/perf/bench_cores.py 262144 2>&1 | tail -10 | tee $OUT/bench_cores_256k.jsonl
