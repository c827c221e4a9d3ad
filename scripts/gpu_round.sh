#!/bin/bash
# One gpurun call: all GPU tests (not -x), smoke, the bench line. Usage: gpurun --timeout 1500 -- bash scripts/gpu_round.sh [n]
set -u
N=${1:-1048576}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench"; timeout 600 python bench.py --particles $N 2>&1 | tail -1 | tee $OUT/bench.json
