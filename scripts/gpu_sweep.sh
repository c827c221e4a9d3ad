#!/bin/bash
# 1-GPU size sweep of the headline kernel (BASELINE configs[4]): gpurun --timeout 400 -- bash scripts/gpu_sweep.sh
set -u
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/bench_sweep_1gpu.jsonl
for N in 262144 524288 1048576 2097152; do
  timeout 200 python bench.py --particles $N --steps 3 --warmup 3 --e2e-steps 1 --cpu-seconds 3 2>&1 | tail -1 >> $OUT/bench_sweep_1gpu.jsonl
done
python - <<'PY'
import json
for l in open("gpurun_out/bench_sweep_1gpu.jsonl"):
    d = json.loads(l)
    print(d["config"]["n_particles"], "%.4g" % d["value"], "frac %.4f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"], d.get("parity", {}).get("vel_err"), d.get("parity", {}).get("grad_err"))
PY
