#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a bench line, the ncu launch list and one full capture of the
# dominant kernel. Everything lands in gpurun_out/. Usage: gpurun --timeout 1500 -- bash scripts/gpu_check.sh [n]
set -u
N=${1:-1048576}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --particles $N 2>&1 | tail -1 | tee $OUT/bench_ref.json
echo "== bench"; timeout 900 python bench.py --particles $N 2>&1 | tail -1 | tee $OUT/bench.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --particles 262144 --no-cpu --e2e-steps 1 > $OUT/bench_under_ncu.log 2>&1
echo "== ncu full capture of the dominant kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pp2_kernel -s 1 -c 1 -f -o $OUT/pp2_full \
    python bench.py --steps 1 --warmup 1 --particles 262144 --no-cpu --e2e-steps 1 > $OUT/ncu_full.log 2>&1
ls -la $OUT
