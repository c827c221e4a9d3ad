#!/bin/bash
# Panel / reflect kernels: every GPU test, then the sphere benchmark at 320 and 5120 panels x 1M points.
# Usage: gpurun --timeout 500 -- bash scripts/gpu_panels.sh
set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee $OUT/pytest_gpu.txt
timeout 200 python tests/perf/bench_panels.py 2 1000000 2>&1 | tail -7 > $OUT/panels_320.jsonl
timeout 200 python tests/perf/bench_panels.py 4 1000000 2>&1 | tail -7 > $OUT/panels_5120.jsonl
python - <<'PY'
import json
for n in (320, 5120):
    for l in open(f"gpurun_out/panels_{n}.jsonl"):
        d = json.loads(l)
        print(n, d["routine"], "%.2f ms" % d["kernel_ms"], d.get("max_rel_err_vs_oracle_sample"), d.get("bit_identical_to_oracle_on_sample"))
PY
