#!/bin/bash
# bench of the 5120-panel sphere + one full ncu capture of pan_pts_kernel<GRAD> (5120 panels x 262144 points).
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python tests/perf/bench_panels.py 4 1000000 2>&1 | tail -7 > $OUT/panels_5120.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/panels_5120.jsonl"):
    d = json.loads(l); print(d["routine"], "%.2f ms" % d["kernel_ms"])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pan_pts_kernel -s 2 -c 1 -f -o $OUT/pan_pts_full \
    python tests/perf/bench_panels.py 4 262144 > $OUT/ncu_pan.log 2>&1
ncu -i $OUT/pan_pts_full.ncu-rep --page raw --csv > $OUT/pan_pts_full_raw.csv 2>/dev/null
ls -la $OUT/pan_pts_full* | head
