#!/bin/bash
# Round-end rehearsal of what the driver runs on a fresh box: the GPU suite, smoke(), both bench arms with the driver's flags
# (--steps 20 --warmup 5) at the default (north-star) size, with wall-clock times. Usage: gpurun --timeout 1800 -- bash scripts/gpu_final.sh
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
t0=$(date +%s)
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/final_pytest.txt
t1=$(date +%s); echo "pytest -m gpu: $((t1-t0)) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/final_smoke.txt
t2=$(date +%s); echo "smoke: $((t2-t1)) s"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/final_bench_reference.json 2> $OUT/final_bench_reference.err
t3=$(date +%s); echo "bench --impl reference: $((t3-t2)) s"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/final_bench.json 2> $OUT/final_bench.err
t4=$(date +%s); echo "bench: $((t4-t3)) s"
python - <<'PY'
import json
for f in ("gpurun_out/final_bench_reference.json","gpurun_out/final_bench.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.4e"%d["value"], d["ms_per_step"], d.get("roofline",{}).get("frac"), d.get("roofline",{}).get("traffic"), d["e2e"]["value"], d.get("parity",{}).get("ok"), d["cpu_baseline"]["cores"])
    except Exception as e: print(f, "failed", e)
PY
tail -2 $OUT/final_bench.err
