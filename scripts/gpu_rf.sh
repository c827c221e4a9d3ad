#!/bin/bash
# register-file bank microbenchmark (+ optional sanitizer pass). Usage: gpurun --timeout 1500 -- bash scripts/gpu_rf.sh [sanitize]
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== rf_banks"; timeout 300 omega3d_b200/csrc/microbench/rf_banks 2>&1 | tee $OUT/rf_banks.txt
if [ "${1:-}" = "sanitize" ]; then bash scripts/gpu_sanitize.sh; fi
