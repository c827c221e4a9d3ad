#!/bin/bash
# Panel / reflect kernels: the GPU tests that touch them, the sphere benchmark at 320 and 5120 panels x 1 M points and at
# 5120 x 262 144 (pooled and per-lane subdivision), one ncu --set full capture of pan_pts_queue_kernel<true> and pts_pan_kernel.
# Usage: gpurun --timeout 900 -- bash scripts/gpu_panels.sh [tag]
set -u
cd "$(dirname "$0")/.."
T=${1:-r2p}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x -k "panel or golden or sphere or dropin or bem or reflect" 2>&1 | tail -5 | tee $OUT/${T}_pytest.txt
for q in queue noqueue; do
  timeout 200 python tests/perf/bench_panels.py 2 1000000 $q 2>&1 | tail -7 > $OUT/${T}_panels_320_$q.jsonl
  timeout 200 python tests/perf/bench_panels.py 4 1000000 $q 2>&1 | tail -7 > $OUT/${T}_panels_5120_$q.jsonl
  timeout 200 python tests/perf/bench_panels.py 4 262144 $q 2>&1 | tail -7 > $OUT/${T}_panels_5120x262k_$q.jsonl
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pan_pts_queue_kernel -s 2 -c 1 -f -o $OUT/${T}_pan_pts_queue \
    python tests/perf/bench_panels.py 4 262144 > $OUT/${T}_ncu_pan.log 2>&1
ncu -i $OUT/${T}_pan_pts_queue.ncu-rep --page raw --csv > $OUT/${T}_pan_pts_queue_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pts_pan_kernel -s 1 -c 1 -f -o $OUT/${T}_pts_pan \
    python tests/perf/bench_panels.py 4 262144 > $OUT/${T}_ncu_ptspan.log 2>&1
ncu -i $OUT/${T}_pts_pan.ncu-rep --page raw --csv > $OUT/${T}_pts_pan_raw.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_panels_*.jsonl")):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        print(f.split("/")[-1], d["routine"], "%.3f ms"%d["kernel_ms"], "err %.2e"%(d.get("max_rel_err_vs_oracle_sample") or 0))
PY
