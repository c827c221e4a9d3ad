#!/bin/bash
# compute-sanitizer over the C ABI at small sizes: memcheck (out-of-bounds / misaligned / leaks of the library's device
# buffers) and initcheck (reads of uninitialised device memory) on the parity tests that touch every kernel.
# Usage: gpurun --timeout 1200 -- bash scripts/gpu_sanitize.sh      -> gpurun_out/sanitize_*.txt
set -u
OUT=gpurun_out; mkdir -p $OUT
SEL="golden or kat or empty or bad_arguments or edge_cases or bit_exact or test_matvec_against_golden or work_pool or pageable or find_vels_with_body or resident_clear_inner or totals_vs_reference"
for tool in memcheck initcheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 540 compute-sanitizer --tool $tool --error-exitcode 66 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py tests/test_gpu_reflect.py tests/test_gpu_convection.py tests/test_gpu_bem.py tests/test_gpu_cores.py tests/test_gpu_sphere.py tests/test_gpu_status.py \
      -q -m gpu -x -k "$SEL" -p no:cacheprovider > $OUT/sanitize_$tool.txt 2>&1
  echo "exit $?" | tee -a $OUT/sanitize_$tool.txt
  grep -E "ERROR SUMMARY|passed|failed|Invalid|Uninitialized" $OUT/sanitize_$tool.txt | tail -8
done
# racecheck (shared-memory hazards) on the particle kernels: the velocity+gradient kernel has no CTA-wide barrier per tile - its
# warps count themselves out of a ring buffer and the last one out refills it through the async proxy - and on the pooled panel kernel
echo "== compute-sanitizer --tool racecheck"
timeout 540 compute-sanitizer --tool racecheck --error-exitcode 66 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden or work_pool" -p no:cacheprovider > $OUT/sanitize_racecheck.txt 2>&1
echo "exit $?" | tee -a $OUT/sanitize_racecheck.txt
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|hazard" $OUT/sanitize_racecheck.txt | tail -8
