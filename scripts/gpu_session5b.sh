#!/bin/bash
# Closing check of session 5: every GPU test at HEAD, smoke, per-core timings, the default bench line.
# Usage: gpurun --timeout 600 -- bash scripts/gpu_session5b.sh
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -q -m gpu --durations=5 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench cores 256K"; timeout 120 python tests/perf/bench_cores.py 262144 2>&1 | tail -10 | tee $OUT/bench_cores_256k.jsonl
echo "== bench cores 1M (vel+grad only is what matters; both printed)"; timeout 200 python tests/perf/bench_cores.py 1048576 2>&1 | tail -10 | tee $OUT/bench_cores_1m.jsonl
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/bench.json
